"""PointNet++ building blocks of the reference's point-cloud branch (`SAModule`, `GlobalSAModule`, `FPModule`,
models/basic_modules.py:66-138) on this package's kernels -- part of the upstream flow producer.
Same constructor arguments, forward signatures and state_dict keys (`conv.local_nn.*`, `nn.*`).  Inference only.

What the reference reaches through torch_cluster / PyG is done by csrc/pointops.cu:
  fps                 farthest point sampling per sample of the batch           (morig_fps)
  radius + PointConv  ball query with at most `max_num_neighbors` hits, then PyG's PointNetConv: a 3-layer MLP on
                      cat[x_j, pos_j - pos_i] per (centre, neighbour) pair and a max over the neighbours.  The first Linear
                      is evaluated per POINT (W [x_j, pos_j - pos_i] + b = (Wx x_j + Wp pos_j + b) - Wp pos_i = Q[j] + P[i]),
                      the pair rows live in a fixed-degree layout [M, K + 1] (unused slots repeat a valid neighbour, which a
                      max does not see), layers 2 / 3 run on the fp32 tile engine with gather / segmented-max epilogues
                      (morig_ball_query, morig_edge_mlp_layer)
  knn_interpolate     k nearest source points + inverse-squared-distance weights (morig_knn_topk, morig_knn_interpolate)
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch
from torch import nn

from . import _lib, engine, packing
from . import train_ops as T
from .basic_modules import MLP, FusedModule  # noqa: F401  (MLP re-exported like the reference module does)


def batch_ptr(batch: torch.Tensor) -> torch.Tensor:
    """int32 [B + 1] segment pointers of a sorted batch vector (one bincount + scan; B is read back once: the result sizes
    of fps / pooling depend on it, as they do in torch_cluster)"""
    counts = torch.bincount(batch)
    ptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=batch.device)
    ptr[1:] = counts.cumsum(0)
    return ptr


def fps(pos: torch.Tensor, batch: torch.Tensor, ratio: float, random_start: bool = True) -> torch.Tensor:
    """`torch_cluster.fps(pos, batch, ratio, random_start)` -> int64 indices, sample by sample, in selection order"""
    lib = _lib.load()
    pos = _lib.require_cuda(pos, "pos")
    ptr = batch_ptr(batch)
    sizes = (ptr[1:] - ptr[:-1]).tolist()                                  # host: output size is data dependent
    m = [int(math.ceil(ratio * n)) for n in sizes]
    out_ptr = torch.tensor([0] + list(torch.tensor(m).cumsum(0).tolist()), dtype=torch.int32, device=pos.device)
    start = None
    if random_start:
        start = torch.tensor([int(torch.randint(0, n, (1,))) for n in sizes], dtype=torch.int32, device=pos.device)
    out = torch.empty(sum(m), dtype=torch.int32, device=pos.device)
    with torch.cuda.device(pos.device):
        _lib.check(lib.morig_fps(pos.data_ptr(), ptr.data_ptr(), out_ptr.data_ptr(), _lib.ptr(start), len(sizes), max(sizes),
                                 out.data_ptr(), _lib.stream_ptr()), "morig_fps")
    return out.long()


def knn(x: torch.Tensor, y: torch.Tensor, k: int, batch_x: torch.Tensor, batch_y: torch.Tensor, cosine: bool = False,
        return_score: bool = False):
    """`torch_cluster.knn(x, y, k, batch_x, batch_y, cosine)` -> [2, k * Ny] (row 0 = y index, row 1 = x index), nearest
    first.  Every segment must hold at least k points."""
    lib = _lib.load()
    x, y = _lib.require_cuda(x, "x"), _lib.require_cuda(y, "y")
    ptr = batch_ptr(batch_x)
    m, d = y.shape
    nbr = torch.empty(m, k, dtype=torch.int32, device=x.device)
    score = torch.empty(m, k, dtype=torch.float32, device=x.device) if return_score else None
    yb = batch_y.to(torch.int32)
    with torch.cuda.device(x.device):
        _lib.check(lib.morig_knn_topk(x.data_ptr(), x.shape[1], ptr.data_ptr(), y.data_ptr(), d, yb.data_ptr(), m, d, k,
                                      1 if cosine else 0, nbr.data_ptr(), _lib.ptr(score), _lib.stream_ptr()), "morig_knn_topk")
    rows = torch.arange(m, device=x.device).repeat_interleave(k)
    assign = torch.stack([rows, nbr.reshape(-1).long()])
    return (assign, score) if return_score else assign


def knn_interpolate(x, pos_x, pos_y, batch_x, batch_y, k: int = 3) -> torch.Tensor:
    """PyG `knn_interpolate`"""
    lib = _lib.load()
    x = _lib.require_cuda(x, "x")
    pos_x, pos_y = _lib.require_cuda(pos_x, "pos_x"), _lib.require_cuda(pos_y, "pos_y")
    ptr = batch_ptr(batch_x)
    m, c = pos_y.shape[0], x.shape[1]
    nbr = torch.empty(m, k, dtype=torch.int32, device=x.device)
    yb = batch_y.to(torch.int32)
    out = torch.empty(m, c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.morig_knn_topk(pos_x.data_ptr(), 3, ptr.data_ptr(), pos_y.data_ptr(), 3, yb.data_ptr(), m, 3, k, 0,
                                      nbr.data_ptr(), 0, _lib.stream_ptr()), "morig_knn_topk")
        _lib.check(lib.morig_knn_interpolate(x.data_ptr(), c, pos_x.data_ptr(), pos_y.data_ptr(), nbr.data_ptr(), m, k, c,
                                             out.data_ptr(), c, _lib.stream_ptr()), "morig_knn_interpolate")
    return out


class DenseStack:
    """eval-mode `MLP` / `Seq(MLP, Linear)` runner: every block is one fused dense layer (Linear -> ReLU -> BatchNorm
    affine); packs are rebuilt when the owning module's weights change (FusedModule.weights_fingerprint)"""

    def __init__(self, owner: FusedModule, name: str, seq: nn.Module):
        self.owner, self.name, self.seq = owner, name, seq

    def _pack(self):
        sd = {"s." + k: v for k, v in self.seq.state_dict(keep_vars=True).items()}
        layers = []

        def walk(prefix, mod):
            if isinstance(mod, nn.Linear):
                layers.append(packing.pack_linear(sd, prefix))
            elif isinstance(mod, nn.Sequential) and len(mod) == 3 and isinstance(mod[0], nn.Linear) and isinstance(mod[2], nn.BatchNorm1d):
                layers.append(packing.pack_mlp_layer(sd, prefix))
            else:
                for i, child in enumerate(mod):
                    walk(f"{prefix}.{i}", child)
        walk("s", self.seq)
        return layers

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        cache = self.owner.__dict__.setdefault("_stacks", {})
        fp = self.owner.weights_fingerprint()
        ent = cache.get(self.name)
        if ent is None or ent[0] != fp:
            with torch.no_grad():
                ent = cache[self.name] = (fp, self._pack())
        for layer in ent[1]:
            m = x.shape[0]
            out = torch.empty(m, layer.N, dtype=torch.float32, device=x.device)
            engine.dense(layer, x, 0, x.stride(0), m, C=out, ldc=layer.N)
            x = out
        return x


class PointConv(nn.Module):
    """parameter container with PyG's key layout (`local_nn.*`); evaluated by SAModule"""

    def __init__(self, local_nn: nn.Module):
        super().__init__()
        self.local_nn = local_nn


class SAModule(FusedModule):
    """`SAModule(ratio, r, nn, max_num_neighbors)` -- models/basic_modules.py:67-89.
    forward(x, pos, batch, random_start=True) -> (x' [M, C'], pos[idx], batch[idx])"""

    def __init__(self, ratio, r, nn, max_num_neighbors):
        super().__init__()
        self.ratio, self.r, self.max_num_neighbors = ratio, r, max_num_neighbors
        self.conv = PointConv(nn)

    def _pack(self):
        sd = {"n." + k: v for k, v in self.conv.local_nn.state_dict(keep_vars=True).items()}
        f64 = packing._f64
        w0, b0 = f64(sd["n.0.0.weight"]), f64(sd["n.0.0.bias"])                  # [h1, C + 3] on cat[x_j, pos_j - pos_i]
        s0, t0 = packing.bn_affine(sd, "n.0.2")
        w1, b1 = f64(sd["n.1.0.weight"]), f64(sd["n.1.0.bias"])
        s1, t1 = packing.bn_affine(sd, "n.1.2")
        w2, b2 = f64(sd["n.2.0.weight"]), f64(sd["n.2.0.bias"])
        s2, t2 = packing.bn_affine(sd, "n.2.2")
        c = w0.shape[1] - 3
        k_src = packing._r4(c + 3)
        q_layer = packing.DenseLayer(W=packing._pack_wt(w0, k_src), K=k_src, N=w0.shape[0], bias=packing._vec(b0))
        p_layer = packing.DenseLayer(W=packing._pack_wt(-w0[:, c:], 4), K=4, N=w0.shape[0])
        l1 = dict(W=packing._pack_wt(w1 * s0.unsqueeze(0)), bias=packing._vec(b1 + w1 @ t0), N=w1.shape[0], K=w1.shape[1])
        l2 = dict(W=packing._pack_wt(w2 * s1.unsqueeze(0)), bias=packing._vec(b2 + w2 @ t1), N=w2.shape[0], K=w2.shape[1],
                  scale=packing._vec(s2), shift=packing._vec(t2))
        return c, q_layer, p_layer, l1, l2

    def forward(self, x: Optional[torch.Tensor], pos: torch.Tensor, batch: torch.Tensor, random_start: bool = True):
        self._guard(pos, batch)
        if self.training:
            raise NotImplementedError("morig_b200.pointnet2: inference only (the training path covers the rigging networks)")
        lib = _lib.load()
        pos = _lib.require_cuda(pos, "pos")
        dev = pos.device
        c, q_layer, p_layer, l1, l2 = self._packed_for("sa", self._pack)
        idx = fps(pos, batch, self.ratio, random_start)
        centres = pos[idx].contiguous()
        cb = batch[idx]
        n, m, K = pos.shape[0], idx.shape[0], self.max_num_neighbors
        ptr = batch_ptr(batch)
        nbr = torch.empty(m, K, dtype=torch.int32, device=dev)
        count = torch.empty(m, dtype=torch.int32, device=dev)
        cb32 = cb.to(torch.int32)
        with torch.cuda.device(dev):
            _lib.check(lib.morig_ball_query(pos.data_ptr(), ptr.data_ptr(), centres.data_ptr(), cb32.data_ptr(), m, float(self.r), K,
                                            nbr.data_ptr(), count.data_ptr(), _lib.stream_ptr()), "morig_ball_query")
        # PyG PointNetConv on the bipartite graph: pairs whose source INDEX equals the centre INDEX are dropped, then the
        # index pairs (i, i) are appended (add_self_loops with num_nodes = min(N, M)); unused slots repeat that pair
        ar = torch.arange(m, dtype=torch.int32, device=dev).unsqueeze(1)
        valid = (torch.arange(K, device=dev).unsqueeze(0) < count.unsqueeze(1)) & (nbr != ar)
        col = torch.cat([torch.where(valid, nbr, ar), ar], dim=1).reshape(-1).contiguous()          # [M * (K + 1)]
        tgt = ar.expand(m, K + 1).reshape(-1).contiguous()
        rowptr = (torch.arange(m + 1, dtype=torch.int32, device=dev) * (K + 1)).contiguous()
        E = m * (K + 1)
        # first Linear per point: Q over all sources on cat[x, pos], P over the centres on pos
        src = torch.zeros(n, q_layer.K, dtype=torch.float32, device=dev)
        if c:
            engine.gather_cols(_lib.require_cuda(x, "x"), x.shape[1], 0, 0, None, c, n, 1, src, q_layer.K, 0)
        engine.gather_cols(pos, 3, 0, 0, None, 3, n, 1, src, q_layer.K, c)
        h1 = q_layer.N
        Q = torch.empty(n, h1, dtype=torch.float32, device=dev)
        engine.dense(q_layer, src, 0, q_layer.K, n, C=Q, ldc=h1)
        cpos = torch.zeros(m, 4, dtype=torch.float32, device=dev)
        engine.gather_cols(centres, 3, 0, 0, None, 3, m, 1, cpos, 4, 0)
        P = torch.empty(m, h1, dtype=torch.float32, device=dev)
        engine.dense(p_layer, cpos, 0, 4, m, C=P, ldc=h1)
        # layers 2 and 3 on the pair rows
        c1 = torch.empty(E, l1["N"], dtype=torch.float32, device=dev)
        out = torch.full((m, l2["N"]), float("-inf"), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            d = _lib.EdgeLayerDesc()
            d.P, d.Q, d.ldpq = P.data_ptr(), Q.data_ptr(), h1
            d.rowptr, d.col, d.tgt, d.n_targets, d.E = rowptr.data_ptr(), col.data_ptr(), tgt.data_ptr(), m, E
            d.W, d.ldw, d.bias = l1["W"].data_ptr(), l1["W"].shape[1], l1["bias"].data_ptr()
            d.C, d.ldc, d.N, d.K = c1.data_ptr(), l1["N"], l1["N"], l1["K"]
            _lib.check(lib.morig_edge_mlp_layer(ctypes.byref(d), _lib.stream_ptr()), "morig_edge_mlp_layer")
            d2 = _lib.EdgeLayerDesc()
            d2.A, d2.lda = c1.data_ptr(), l1["N"]
            d2.rowptr, d2.col, d2.tgt, d2.n_targets, d2.E = rowptr.data_ptr(), col.data_ptr(), tgt.data_ptr(), m, E
            d2.W, d2.ldw, d2.bias = l2["W"].data_ptr(), l2["W"].shape[1], l2["bias"].data_ptr()
            d2.scale, d2.shift = l2["scale"].data_ptr(), l2["shift"].data_ptr()
            d2.out, d2.ldo, d2.N, d2.K = out.data_ptr(), l2["N"], l2["N"], l2["K"]
            _lib.check(lib.morig_edge_mlp_layer(ctypes.byref(d2), _lib.stream_ptr()), "morig_edge_mlp_layer")
        return out, centres, cb


class GlobalSAModule(FusedModule):
    """`GlobalSAModule(nn)` -- models/basic_modules.py:118-128"""

    def __init__(self, nn):
        super().__init__()
        self.nn = nn

    def forward(self, x, pos, batch):
        self._guard(x, pos, batch)
        if self.training:
            raise NotImplementedError("morig_b200.pointnet2: inference only")
        h = DenseStack(self, "nn", self.nn)(T.concat_cols([x, pos]))
        b32 = batch.to(torch.int32)
        n_seg = int(batch[-1].item()) + 1
        out, _ = T.segmax_fwd(h, T.seg_ptr(b32, n_seg), n_seg)                 # global_max_pool
        return out, pos.new_zeros((n_seg, 3)), torch.arange(n_seg, device=batch.device)


class FPModule(FusedModule):
    """`FPModule(k, nn)` -- models/basic_modules.py:130-141"""

    def __init__(self, k, nn):
        super().__init__()
        self.k = k
        self.nn = nn

    def forward(self, x, pos, batch, x_skip, pos_skip, batch_skip):
        self._guard(x, pos, batch)
        if self.training:
            raise NotImplementedError("morig_b200.pointnet2: inference only")
        y = knn_interpolate(x, pos, pos_skip, batch, batch_skip, k=self.k)
        if x_skip is not None:
            y = T.concat_cols([y, x_skip])
        return DenseStack(self, "nn", self.nn)(y), pos_skip, batch_skip
