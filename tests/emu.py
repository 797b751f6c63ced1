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


def fill_many(tensors, value):
    for t in tensors:
        t.fill_(value)


def fill_cut(entries, n, n_frames, value):
    """emulation of morig_fill_cut_f32 (include/morig_b200.h): the vertices whose CSR segment straddles a multiple of 32 slots"""
    for g, out, ld, col0, ncols in entries:
        rp = g.rowptr.long()
        a, b = rp[:-1], rp[1:]
        cut = (b > a) & ((a >> 5) != ((b - 1) >> 5))
        rows = _view(out, 0, n * n_frames, ld, ld).reshape(n_frames, n, ld)
        rows[:, cut, col0:col0 + ncols] = value


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
    # float64 is let through: the training host logic is also checked in double precision (exactly, no rounding chaos)
    if t.dtype != dtype and not (dtype == torch.float32 and t.dtype == torch.float64):
        raise TypeError(f"{name}: {t.dtype}")
    return t.contiguous()


def install(monkeypatch):
    for name in ("graph_prep", "dense", "edgeconv", "edgeconv_batch", "fill", "fill_many", "fill_cut", "gather_cols", "row_normalize", "temporal_attn",
                 "frame_reduce"):
        monkeypatch.setattr(engine, name, globals()[name])
    monkeypatch.setattr(_lib, "require_cuda", _require)

    def guard(self, *tensors):
        pass
    monkeypatch.setattr(basic_modules.FusedModule, "_guard", guard)


# ---- training path: torch-CPU restatements of morig_b200/train_ops.py (semantics of include/morig_b200.h) -----------------
def _t_linear_fwd(x, w, b, relu=False):
    y = x @ w.t()
    if b is not None:
        y = y + b
    return torch.relu(y) if relu else y


def _t_matmul_nn(dy, w):
    return dy @ w


def _t_wgrad(dy, x, want_bias):
    return (dy.double().t() @ x.double()).to(x.dtype), (dy.double().sum(0).to(x.dtype) if want_bias else None)


def _t_bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum, eps=1e-5):
    R = x.shape[0]
    xd = x.double()
    mu = xd.mean(0)
    var = (xd * xd).mean(0) - mu * mu
    var = var.clamp(min=0)
    invstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * invstd
    shift = beta.double() - mu * scale
    if running_mean is not None:
        running_mean.copy_(((1 - momentum) * running_mean.double() + momentum * mu).to(x.dtype))
    if running_var is not None:
        unb = var * (R / (R - 1)) if R > 1 else var
        running_var.copy_(((1 - momentum) * running_var.double() + momentum * unb).to(x.dtype))
    return x * scale.to(x.dtype) + shift.to(x.dtype), mu.to(x.dtype), invstd.to(x.dtype)


def _t_bn_relu_bwd(dy, x, gamma, mean, invstd, relu):
    R = x.shape[0]
    xhat = (x - mean) * invstd
    s0 = dy.double().sum(0)
    s1 = (dy.double() * x.double()).sum(0)
    dgamma = invstd.double() * (s1 - mean.double() * s0)
    a = (gamma.double() * invstd.double()).to(x.dtype)
    g = a * (dy - (s0 / R).to(x.dtype) - xhat * (dgamma / R).to(x.dtype))
    if relu:
        g = torch.where(x > 0, g, torch.zeros_like(g))
    return g, dgamma.to(x.dtype), s0.to(x.dtype)


def _t_edge_gather_relu(P, Q, g):
    e = g.e_real
    return torch.relu(P[g.tgt[:e].long()] + Q[g.col[:e].long()])


def _t_edge_gather_relu_bwd(dh, h, g):
    e, C = g.e_real, h.shape[1]
    dz = torch.where(h > 0, dh, torch.zeros_like(dh))
    dpq = torch.zeros(g.n, 2 * C, dtype=h.dtype)
    dpq[:, :C].index_add_(0, g.tgt[:e].long(), dz)
    dpq[:, C:].index_add_(0, g.col[:e].long(), dz)
    return dpq


def _t_segmax_fwd(y, ptr, S):
    C = y.shape[1]
    out = torch.zeros(S, C, dtype=y.dtype)
    arg = torch.full((S, C), -1, dtype=torch.int32)
    p = ptr.tolist()
    for s in range(S):
        if p[s + 1] > p[s]:
            seg = y[p[s]:p[s + 1]]
            m = seg.max(0).values
            first = (seg == m).to(torch.uint8).argmax(0)  # first maximal row
            out[s] = m
            arg[s] = (first + p[s]).to(torch.int32)
    return out, arg


def _t_segmax_bwd(dout, arg, R):
    dy = torch.zeros(R, arg.shape[1], dtype=dout.dtype)
    ok = arg >= 0
    cols = torch.arange(arg.shape[1]).expand_as(arg)
    dy[arg[ok].long(), cols[ok]] = dout[ok]
    return dy


def _t_seg_ptr(keys32, S):
    counts = torch.bincount(keys32.long(), minlength=S)
    return torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).to(torch.int32)


def _t_row_gather(src, idx32):
    return src[idx32.long()]


def _t_seg_sum(x, ptr, S):
    p = ptr.tolist()
    return torch.stack([x[p[s]:p[s + 1]].double().sum(0).to(x.dtype) for s in range(S)])


def _t_normalize_fwd(x):
    return x / x.norm(dim=1, keepdim=True).clamp(min=1e-12)


def _t_normalize_bwd(x, dy):
    n = x.norm(dim=1, keepdim=True).clamp(min=1e-12)
    return dy / n - x * ((x * dy).sum(1, keepdim=True) / n ** 3)


def _t_attn_cls_fwd(q0, kc, vc, Kx, Vx, d):
    N, T, HD = Kx.shape
    heads = HD // d
    q = q0.reshape(heads, d)
    k = torch.cat([kc.reshape(1, 1, HD).expand(N, 1, HD), Kx], 1).reshape(N, T + 1, heads, d)
    v = torch.cat([vc.reshape(1, 1, HD).expand(N, 1, HD), Vx], 1).reshape(N, T + 1, heads, d)
    logits = torch.einsum("hd,nthd->nht", q, k) / d ** 0.5
    att = torch.softmax(logits, dim=2)
    out = torch.einsum("nht,nthd->nhd", att, v).reshape(N, HD)
    return out, att.contiguous()


def _t_attn_cls_bwd(q0, kc, vc, Kx, Vx, att, dout, d):
    N, T, HD = Kx.shape
    heads = HD // d
    q = q0.reshape(heads, d)
    k = torch.cat([kc.reshape(1, 1, HD).expand(N, 1, HD), Kx], 1).reshape(N, T + 1, heads, d)
    v = torch.cat([vc.reshape(1, 1, HD).expand(N, 1, HD), Vx], 1).reshape(N, T + 1, heads, d)
    g = dout.reshape(N, heads, d)
    da = torch.einsum("nhd,nthd->nht", g, v)
    dl = att * (da - (att * da).sum(2, keepdim=True)) / d ** 0.5
    dv = torch.einsum("nht,nhd->nthd", att, g)
    dk = torch.einsum("nht,hd->nthd", dl, q)
    dq = torch.einsum("nht,nthd->hd", dl, k)
    return (dq.reshape(HD), dk[:, 0].sum(0).reshape(HD), dv[:, 0].sum(0).reshape(HD),
            dk[:, 1:].reshape(N, T, HD).contiguous(), dv[:, 1:].reshape(N, T, HD).contiguous())


def _t_scale_cols(x, factor):
    return x * factor


def _t_concat_cols(xs, pad_to=1):
    y = torch.cat(list(xs), dim=1)
    extra = (-y.shape[1]) % pad_to
    return y if extra == 0 else torch.cat([y, y.new_zeros(y.shape[0], extra)], dim=1)


def install_train(monkeypatch):
    """swap the launch helpers of morig_b200.train_ops for the restatements above (host-logic tests of the training path)"""
    from morig_b200 import train_ops
    for name in ("linear_fwd", "matmul_nn", "wgrad", "bn_train_fwd", "bn_relu_bwd", "edge_gather_relu", "edge_gather_relu_bwd",
                 "segmax_fwd", "segmax_bwd", "seg_ptr", "row_gather", "seg_sum", "normalize_fwd", "normalize_bwd",
                 "attn_cls_fwd", "attn_cls_bwd", "concat_cols", "scale_cols"):
        monkeypatch.setattr(train_ops, name, globals()["_t_" + name])
