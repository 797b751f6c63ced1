"""TEST INFRASTRUCTURE — CPU emulation of the C-ABI entry points, for checking HOST logic only.

The build container has no GPU.  To still exercise `morig_b200`'s host side there (weight folding
and packing, buffer layouts, launch sequences, cache invalidation), the `emulated` fixture swaps the
thin launch helpers of `morig_b200.engine` for torch-CPU restatements of the semantics documented in
include/morig_b200.h.  This lives under tests/ and is never importable from the product package;
the GPU tests (`-m gpu`) run the real library and do not use it.
"""
from __future__ import annotations

import numpy as np
import torch

from morig_b200 import _lib, basic_modules, engine
from oracle import graph_port


def _view(t: torch.Tensor, off: int, rows: int, cols: int, ld: int) -> torch.Tensor:
    flat = t.reshape(-1)
    return torch.as_strided(flat, (rows, cols), (ld, 1), storage_offset=flat.storage_offset() + off)


def graph_prep(edge_index, n, overlap=False):
    rowptr, col = graph_port.csr_by_target(edge_index.numpy(), n)
    tgt = np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr))
    e = edge_index.shape[1]
    pad = e + n - col.shape[0]
    col = np.concatenate([col, np.full(pad, -1, np.int32)])
    tgt = np.concatenate([tgt, np.full(pad, -1, np.int32)])
    return engine.Graph(rowptr=torch.from_numpy(rowptr), col=torch.from_numpy(col), tgt=torch.from_numpy(tgt),
                        n=n, e_max=e + n)


def _groups(M, binfo, n_vtx):
    r = torch.arange(M)
    return (r // n_vtx) * binfo.n_graphs + binfo.batch32.long()[r % n_vtx]


def dense(layer, A, a_off, lda, M, *, K=None, C=None, c_off=0, ldc=0, pool=None, rowbias=None, binfo=None, n_vtx=0):
    K = layer.K if K is None else K
    a = _view(A, a_off, M, K, lda)
    v = a @ layer.W[:K, :layer.N]
    if layer.bias is not None:
        v = v + layer.bias
    if rowbias is not None:
        v = v + rowbias[_groups(M, binfo, n_vtx)][:, :layer.N]
    if layer.relu:
        v = torch.relu(v)
    if layer.scale is not None:
        v = v * layer.scale + layer.shift
    if C is not None:
        _view(C, c_off, M, layer.N, ldc).copy_(v)
    if pool is not None:
        g = _groups(M, binfo, n_vtx)
        pool.copy_(pool.scatter_reduce(0, g.unsqueeze(1).expand_as(v), v, reduce="amax", include_self=True))


def edgeconv(br, pq, ldpq, p_off, q_off, g, n_frames, out, ldo, out_off, out_repeat=1):
    n, H = g.n, br.H
    e_real = int(g.rowptr[n])
    i, j = g.tgt[:e_real].long(), g.col[:e_real].long()
    for f in range(n_frames):
        P = _view(pq, f * n * ldpq + p_off, n, H, ldpq)
        Q = _view(pq, f * n * ldpq + q_off, n, H, ldpq)
        h = torch.relu(P[i] + Q[j])
        z = torch.relu(h @ br.W1[:H, :H] + br.b1) * br.scale + br.shift
        m = torch.full((n, H), float("-inf")).scatter_reduce(0, i.unsqueeze(1).expand_as(z), z, reduce="amax",
                                                              include_self=True)
        for r in range(out_repeat):
            _view(out, (f + r) * n * ldo + out_off, n, H, ldo).copy_(m)


def edgeconv_batch(items, g, n_frames, out_repeat=1):
    for br, pq, ldpq, p_off, q_off, out, ldo, out_off in items:
        edgeconv(br, pq, ldpq, p_off, q_off, g, n_frames, out, ldo, out_off, out_repeat)


def fill(t, value):
    t.fill_(value)


def gather_cols(src, lds, src_off, frame_stride, cols, c, n, n_frames, dst, ldd, dst_off):
    cc = torch.arange(c) if cols is None else cols.long()
    for f in range(n_frames):
        s = _view(src, 0, n, lds, lds)[:, src_off + f * frame_stride + cc]
        _view(dst, f * n * ldd + dst_off, n, c, ldd).copy_(s)


def row_normalize(x, ldx, rows, c, dst2=None, n=0, n_frames=0):
    v = _view(x, 0, rows, c, ldx)
    v.copy_(torch.nn.functional.normalize(v, dim=1))
    if dst2 is not None:
        dst2.copy_(v.reshape(n_frames, n, c).permute(1, 0, 2))


def temporal_attn(pk, x, out):
    n, t, c = x.shape
    res = torch.zeros(n, pk.D)
    for h in range(pk.heads):
        logits = torch.cat([pk.l0[h].expand(n, 1), x @ pk.u[h]], dim=1)
        a = torch.softmax(logits, dim=1)
        y = (a[:, 1:, None] * x).sum(1)
        res = res + y @ pk.Mv[h].t() + a[:, :1] * pk.c0[h]
    out.copy_(res)


def frame_reduce(x, mode, out):
    out.copy_(x.mean(1) if mode == "mean" else x.max(1)[0])


def _require(t, name, dtype=torch.float32):
    if t.dtype != dtype:
        raise TypeError(f"{name}: {t.dtype}")
    return t.contiguous()


def install(monkeypatch):
    for name in ("graph_prep", "dense", "edgeconv", "edgeconv_batch", "fill", "gather_cols", "row_normalize", "temporal_attn",
                 "frame_reduce"):
        monkeypatch.setattr(engine, name, globals()[name])
    monkeypatch.setattr(_lib, "require_cuda", _require)

    def guard(self, *tensors):
        if self.training:
            raise NotImplementedError("train mode")
    monkeypatch.setattr(basic_modules.FusedModule, "_guard", guard)
