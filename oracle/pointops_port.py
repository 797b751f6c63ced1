"""TEST INFRASTRUCTURE (oracle): torch-CPU restatements of the third-party point-cloud operators the reference's CorrNet /
DeformNet reach (models/corrnet.py, models/deformnet.py, models/basic_modules.py:66-138):

    torch_cluster 1.6.0   fps(pos, batch, ratio, random_start), knn(x, y, k, batch_x, batch_y, cosine)
    PyG 2.0.4             knn_interpolate, global_max_pool, PointConv (= PointNetConv)

Their sources are not in /root/reference ("restated, unpinned", like oracle/pyg_shim.py); `install()` registers them in the
shim modules so that the reference's own, unmodified CorrNet / DeformNet run on the CPU (where the reference itself
switches to its `radius_cpu`, models/basic_modules.py:9-30, and to the matmul / max form of the 1-NN, models/corrnet.py:66-73).
"""
from __future__ import annotations

import math
import sys

import torch

from . import pyg_shim


def segments(batch: torch.Tensor):
    counts = torch.bincount(batch)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    return ptr.tolist()


def fps(pos, batch=None, ratio=0.5, random_start=True):
    """torch_cluster.fps: per segment ceil(ratio * n) samples, first = the segment's first point (random_start=False),
    then repeatedly the point farthest (squared Euclidean, fp32) from the chosen set; ties -> lower index.
    Returns global indices, segment by segment, in selection order."""
    if batch is None:
        batch = torch.zeros(pos.shape[0], dtype=torch.long)
    ptr = segments(batch)
    out = []
    for b in range(len(ptr) - 1):
        p = pos[ptr[b]:ptr[b + 1]]
        n = p.shape[0]
        m = int(math.ceil(ratio * n))
        cur = int(torch.randint(0, n, (1,))) if random_start else 0
        dmin = torch.full((n,), float("inf"))
        for _ in range(m):
            out.append(ptr[b] + cur)
            d = ((p - p[cur]) ** 2)
            d = (d[:, 0] + d[:, 1]) + d[:, 2]
            dmin = torch.minimum(dmin, d)
            cur = int(torch.argmax(dmin))          # first maximum
    return torch.tensor(out, dtype=torch.long)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32):
    """torch_cluster.radius (CUDA semantics): for every y_i the first `max_num_neighbors` points x_j (index order) of the
    same segment with |x_j - y_i|^2 < r^2.  Returns [2, E]: row 0 = y index, row 1 = x index."""
    rows, cols = [], []
    px = segments(batch_x)
    for i in range(y.shape[0]):
        b = int(batch_y[i])
        seg = x[px[b]:px[b + 1]]
        d = (seg - y[i]) ** 2
        d = (d[:, 0] + d[:, 1]) + d[:, 2]
        hit = torch.nonzero(d < r * r).squeeze(1)[:max_num_neighbors] + px[b]
        rows.append(torch.full((hit.numel(),), i, dtype=torch.long))
        cols.append(hit)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def knn(x, y, k, batch_x=None, batch_y=None, cosine=False, num_workers=1):
    """torch_cluster.knn: for every y_i its k nearest x_j of the same segment (Euclidean, or cosine distance
    1 - x.y / (|x| |y|)), nearest first; ties -> lower index.  Returns [2, k * Ny]: row 0 = y index, row 1 = x index."""
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.long)
    if batch_y is None:
        batch_y = torch.zeros(y.shape[0], dtype=torch.long)
    px = segments(batch_x)
    rows, cols = [], []
    for b in range(len(px) - 1):
        xs = x[px[b]:px[b + 1]]
        ids = torch.nonzero(batch_y == b).squeeze(1)
        ys = y[ids]
        if cosine:
            sim = (ys @ xs.t()) / (ys.norm(dim=1, keepdim=True).clamp(min=1e-30) * xs.norm(dim=1).clamp(min=1e-30))
            key = -sim
        else:
            key = ((ys[:, None, :] - xs[None, :, :]) ** 2).sum(-1)
        order = torch.sort(key, dim=1, stable=True).indices[:, :k]
        rows.append(ids.repeat_interleave(order.shape[1]))
        cols.append(order.reshape(-1) + px[b])
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3, num_workers=1):
    """PyG knn_interpolate: inverse squared distance weights over the k nearest source points"""
    with torch.no_grad():
        assign = knn(pos_x, pos_y, k, batch_x, batch_y)
        y_idx, x_idx = assign[0], assign[1]
        diff = pos_x[x_idx] - pos_y[y_idx]
        w = 1.0 / torch.clamp((diff * diff).sum(dim=-1, keepdim=True), min=1e-16)
    num = torch.zeros(pos_y.shape[0], x.shape[1], dtype=x.dtype).index_add_(0, y_idx, x[x_idx] * w)
    den = torch.zeros(pos_y.shape[0], 1, dtype=x.dtype).index_add_(0, y_idx, w)
    return num / den


def global_max_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    out = torch.full((n, x.shape[1]), torch.finfo(x.dtype).min, dtype=x.dtype)
    return out.scatter_reduce(0, batch.unsqueeze(1).expand_as(x), x, reduce="amax", include_self=True)


class PointConv(pyg_shim.MessagePassing):
    """PyG 2.0.4 PointNetConv(local_nn, global_nn=None, add_self_loops=True), aggr='max'.  For the bipartite call of
    SAModule (models/basic_modules.py:88-91) the release first drops edges whose source and target INDEX coincide and then
    appends the pairs (i, i) for i < min(#sources, #targets) -- index pairs, not geometric self loops."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops=True, **kwargs):
        super().__init__(aggr="max")
        self.local_nn, self.global_nn, self.add_self_loops = local_nn, global_nn, add_self_loops

    def forward(self, x, pos, edge_index):
        if not isinstance(x, tuple):
            x = (x, None)
        if torch.is_tensor(pos):
            pos = (pos, pos)
        if self.add_self_loops:
            edge_index, _ = pyg_shim.remove_self_loops(edge_index)
            edge_index, _ = pyg_shim.add_self_loops(edge_index, num_nodes=min(pos[0].size(0), pos[1].size(0)))
        src, dst = edge_index[0], edge_index[1]
        msg = pos[0].index_select(0, src) - pos[1].index_select(0, dst)
        if x[0] is not None:
            msg = torch.cat([x[0].index_select(0, src), msg], dim=1)
        if self.local_nn is not None:
            msg = self.local_nn(msg)
        out, _ = pyg_shim.scatter_max(msg, dst, dim=0, dim_size=pos[1].size(0))
        return out if self.global_nn is None else self.global_nn(out)


def install() -> None:
    """register the restatements in the shim modules (idempotent); the reference's `models` package must be imported
    AFTER this call for its `from torch_geometric.nn import ...` lines to pick them up"""
    pyg_shim.install()
    tgnn = sys.modules["torch_geometric.nn"]
    for name, fn in (("fps", fps), ("radius", radius), ("knn", knn), ("knn_interpolate", knn_interpolate),
                     ("global_max_pool", global_max_pool), ("PointConv", PointConv)):
        setattr(tgnn, name, fn)
    tc = sys.modules["torch_cluster"]
    tc.fps, tc.knn, tc.radius = fps, knn, radius
    # modules of the reference that were imported before keep their own references: patch those too
    for modname in ("models.basic_modules", "models.corrnet", "models.deformnet"):
        m = sys.modules.get(modname)
        if m is not None:
            for name, fn in (("fps", fps), ("radius", radius), ("knn", knn), ("knn_interpolate", knn_interpolate),
                             ("global_max_pool", global_max_pool), ("PointConv", PointConv)):
                if hasattr(m, name):
                    setattr(m, name, fn)
