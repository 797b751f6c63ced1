"""Seeded synthetic rigging batches shaped like the reference's `RigDataset` items.

The reference ships no data (datasets are external downloads), so every parity test and
bench line in this repo runs on the synthetic meshes specified in SURVEY.md §8(d):

* a jittered torus grid with `n_u * n_v = N` vertices, normalised like
  `data_proc/common_ops.py:123-138` (largest bbox side = 1, pivot at bottom centre);
* topological edges: every grid quad split into two triangles (valence 6), both directions,
  rows `[v, n]` in the convention of `get_tpl_edges` (`data_proc/common_ops.py:15-32`);
* "geodesic" edges: 15 nearest neighbours (self excluded) by fp32 squared distance, ties to
  the lower index, rows `[i, nbr]` as `get_geo_edges` (`data_proc/common_ops.py:214-226`);
* self loops appended the way the dataset does (`datasets/dataset_rig.py:121-122`);
* `flow [N, 15]`: five key-frames of piece-wise rigid motion plus noise
  (`gt_flow`/`pred_flow`, `datasets/dataset_rig.py:104-115`);
* `skin_input [N, 160]`: 20 nearest bones x (6 coords, 1/(D+1e-10), is-leaf)
  (`datasets/dataset_rig.py:50-64`).

Batches are collated PyG-style (vertex-concatenated, edge indices offset, `batch` vector).
Everything here is host-side numpy/torch-CPU; it is input generation, not the measured path.
"""
from __future__ import annotations

import types

import numpy as np
import torch

GRID = {1024: (32, 32), 2048: (64, 32), 4096: (64, 64), 8192: (128, 64), 16384: (128, 128)}


def _grid_shape(n_vtx: int):
    if n_vtx in GRID:
        return GRID[n_vtx]
    nu = int(np.sqrt(n_vtx))
    while n_vtx % nu:
        nu -= 1
    nv = n_vtx // nu
    if nu < 3 or nv < 3:
        raise ValueError(f"cannot build a torus grid with {n_vtx} vertices")
    return nv, nu


def torus_vertices(n_vtx: int, rng: np.random.Generator) -> np.ndarray:
    nu, nv = _grid_shape(n_vtx)
    u = (np.arange(nu, dtype=np.float64) / nu) * 2 * np.pi
    v = (np.arange(nv, dtype=np.float64) / nv) * 2 * np.pi
    uu, vv = np.meshgrid(u, v, indexing="ij")
    R, r = 0.35, 0.12
    x = (R + r * np.cos(vv)) * np.cos(uu)
    y = r * np.sin(vv)
    z = (R + r * np.cos(vv)) * np.sin(uu)
    p = np.stack([x, y, z], -1).reshape(-1, 3)
    p = p + rng.normal(0.0, 1e-3, p.shape)
    lo, hi = p.min(0), p.max(0)
    pivot = np.array([(lo[0] + hi[0]) / 2, lo[1], (lo[2] + hi[2]) / 2])
    p = (p - pivot) / (hi - lo).max()
    return p.astype(np.float32)


def torus_tpl_edges(n_vtx: int) -> np.ndarray:
    """[2, 6N] int64, rows (source v, neighbour n); six neighbours of the split-quad grid."""
    nu, nv = _grid_shape(n_vtx)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    vid = (iu * nv + iv).reshape(-1)
    nbrs = []
    for du, dv in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1)):
        nbrs.append((((iu + du) % nu) * nv + ((iv + dv) % nv)).reshape(-1))
    nbr = np.stack(nbrs, 1)                      # [N, 6]
    nbr.sort(axis=1)                             # set() order of the reference is ascending for small ints
    src = np.repeat(vid, 6)
    return np.stack([src, nbr.reshape(-1)], 0).astype(np.int64)


def knn_edges(pos: np.ndarray, k: int = 15) -> np.ndarray:
    """[2, kN] int64 rows (i, nbr): k nearest by fp32 ((a-b)^2).sum(), self excluded, ties -> lower index."""
    p = torch.from_numpy(pos)
    n = p.shape[0]
    k = min(k, n - 1)                            # tiny meshes: every other vertex
    out = np.empty((n, k), dtype=np.int64)
    step = 1024
    for s in range(0, n, step):
        d = ((p[s:s + step, None, :] - p[None, :, :]) ** 2).sum(-1)       # fp32, same expression as the test oracle
        d[torch.arange(d.shape[0]), torch.arange(s, s + d.shape[0])] = float("inf")
        # stable sort == ties to the lower index
        idx = torch.sort(d, dim=1, stable=True).indices[:, :k]
        out[s:s + step] = idx.numpy()
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    return np.stack([src, out.reshape(-1)], 0)


def add_self_loops_np(ei: np.ndarray, n: int) -> np.ndarray:
    loop = np.arange(n, dtype=np.int64)
    return np.concatenate([ei, np.stack([loop, loop], 0)], 1)


def rigid_flow(pos: np.ndarray, rng: np.random.Generator, n_key: int = 5, n_seg: int = 4) -> np.ndarray:
    """Piece-wise rigid key-frame displacement [N, 3*n_key] + N(0, 0.005) noise."""
    ang = np.arctan2(pos[:, 2], pos[:, 0])
    seg = np.minimum(((ang + np.pi) / (2 * np.pi) * n_seg).astype(np.int64), n_seg - 1)
    theta = rng.uniform(-0.3, 0.3, n_seg)
    axis = np.array([0.0, 1.0, 0.0])
    flows = []
    for t in range(1, n_key + 1):
        vt = pos.astype(np.float64).copy()
        for s in range(n_seg):
            m = seg == s
            a0 = -np.pi + s * 2 * np.pi / n_seg
            joint = np.array([0.35 * np.cos(a0), pos[:, 1].mean(), 0.35 * np.sin(a0)])
            th = theta[s] * t / n_key
            c, sn = np.cos(th), np.sin(th)
            K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            Rm = np.eye(3) + sn * K + (1 - c) * (K @ K)
            vt[m] = (vt[m] - joint) @ Rm.T + joint
        flows.append(vt - pos)
    f = np.concatenate(flows, 1) + rng.normal(0.0, 0.005, (pos.shape[0], 3 * n_key))
    return f.astype(np.float32)


def _pts2segs(pts: np.ndarray, segs: np.ndarray) -> np.ndarray:
    a, b = segs[:, 0:3], segs[:, 3:6]
    ab = b - a
    l2 = (ab ** 2).sum(1)
    t = ((pts[:, None, :] - a[None]) * ab[None]).sum(-1) / (l2[None] + 1e-6)
    t = np.clip(t, 0, 1)
    proj = a[None] + t[..., None] * ab[None]
    return np.sqrt(((proj - pts[:, None, :]) ** 2).sum(-1))


def skin_input(pos: np.ndarray, n_joint: int = 32, n_near: int = 20) -> np.ndarray:
    """[N, 8*n_near] per-vertex nearest-bone descriptors (6 bone coords, 1/(D+1e-10), leaf flag)."""
    a = np.arange(n_joint) / n_joint * 2 * np.pi
    cy = float(pos[:, 1].mean())
    sc = 0.35 / 0.94                                  # centre circle after bbox normalisation
    j = np.stack([sc * np.cos(a), np.full_like(a, cy), sc * np.sin(a)], 1)
    bones = np.concatenate([j[:-1], j[1:]], 1)        # chain: 31 bones
    leaf = np.concatenate([j[-1], j[-1] + 0.02 * (j[-1] - j[-2]) / np.linalg.norm(j[-1] - j[-2])])[None]
    bones = np.concatenate([bones, leaf], 0)          # + 1 leaf bone
    isleaf = np.zeros(len(bones)); isleaf[-1] = 1.0
    d = _pts2segs(pos.astype(np.float64), bones)
    order = np.argsort(d, axis=1, kind="stable")[:, :n_near]
    dd = np.take_along_axis(d, order, 1)
    feat = np.concatenate([bones[order], (1.0 / (dd + 1e-10))[..., None], isleaf[order][..., None]], -1)
    feat[..., 6] = np.minimum(feat[..., 6], 1e4)      # keep fp32-friendly magnitudes
    return feat.reshape(pos.shape[0], -1).astype(np.float32)


def make_mesh(n_vtx: int, seed: int, with_skin: bool = False) -> dict:
    rng = np.random.default_rng(seed)
    pos = torus_vertices(n_vtx, rng)
    tpl = add_self_loops_np(torus_tpl_edges(n_vtx), n_vtx)
    geo = add_self_loops_np(knn_edges(pos, 15), n_vtx)
    out = dict(pos=pos, tpl_edge_index=tpl, geo_edge_index=geo, flow=rigid_flow(pos, rng))
    if with_skin:
        out["skin_input"] = skin_input(pos)
    return out


class Batch(types.SimpleNamespace):
    """Attribute bag with `.to(device)`; stands in for a PyG `Batch` (only attribute access is used
    by the reference forward, `models/rignet.py:83,86,159,174-179`)."""

    def to(self, device, non_blocking: bool = False):
        kw = {k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
              for k, v in self.__dict__.items()}
        return Batch(**kw)

    def pin_memory(self):
        kw = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}
        return Batch(**kw)


def collate(meshes: list) -> Batch:
    """PyG-style collation (`datasets/dataset_pose.py:21-25` increments edge indices by #vertices)."""
    off = 0
    pos, flow, tpl, geo, bat, skin = [], [], [], [], [], []
    for b, m in enumerate(meshes):
        n = m["pos"].shape[0]
        pos.append(m["pos"]); flow.append(m["flow"])
        tpl.append(m["tpl_edge_index"] + off); geo.append(m["geo_edge_index"] + off)
        bat.append(np.full(n, b, dtype=np.int64))
        if "skin_input" in m:
            skin.append(m["skin_input"])
        off += n
    data = Batch(pos=torch.from_numpy(np.concatenate(pos)),
                 tpl_edge_index=torch.from_numpy(np.concatenate(tpl, 1)),
                 geo_edge_index=torch.from_numpy(np.concatenate(geo, 1)),
                 batch=torch.from_numpy(np.concatenate(bat)),
                 pred_flow=torch.from_numpy(np.concatenate(flow)),
                 num_graphs=len(meshes))                   # PyG `Batch.num_graphs`
    if skin:
        data.skin_input = torch.from_numpy(np.concatenate(skin))
    return data


_MESH_CACHE: dict = {}


def make_batch(n_graphs: int, n_vtx: int, seed: int = 0, with_skin: bool = False) -> Batch:
    meshes = []
    for b in range(n_graphs):
        key = (n_vtx, seed + b, with_skin)
        if key not in _MESH_CACHE:
            _MESH_CACHE[key] = make_mesh(n_vtx, seed + b, with_skin)
        meshes.append(_MESH_CACHE[key])
    return collate(meshes)


def make_deform_batch(n_graphs: int, n_vtx: int, n_pts: int, seed: int = 0) -> Batch:
    """synthetic input of the upstream flow producer (CorrNet / DeformNet, models/deformnet.py:42): a mesh batch
    (`vtx`, edge lists, `vtx_batch`) plus an observed partial point cloud per mesh (`pts`, `pts_batch`): the key-frame-1
    positions of the vertices on the +z side, sub-sampled to n_pts points with N(0, 0.002) noise"""
    meshes, pts, pb = [], [], []
    for b in range(n_graphs):
        m = make_mesh(n_vtx, seed + b)
        meshes.append(m)
        rng = np.random.default_rng(1000 + seed + b)
        moved = m["pos"] + m["flow"][:, 0:3]
        order = np.argsort(-moved[:, 2], kind="stable")                  # most "visible" (largest z) first
        pick = np.sort(order[: max(n_pts, 1)]) if n_pts <= n_vtx else np.sort(rng.integers(0, n_vtx, n_pts))
        cloud = moved[pick] + rng.normal(0.0, 0.002, (len(pick), 3))
        pts.append(cloud.astype(np.float32))
        pb.append(np.full(len(pick), b, dtype=np.int64))
    d = collate(meshes)
    return Batch(vtx=d.pos, tpl_edge_index=d.tpl_edge_index, geo_edge_index=d.geo_edge_index, vtx_batch=d.batch,
                 pts=torch.from_numpy(np.concatenate(pts)), pts_batch=torch.from_numpy(np.concatenate(pb)),
                 num_graphs=n_graphs)


def randomize_bn_(model: torch.nn.Module, seed: int = 0) -> None:
    """Make eval-mode BatchNorm non-trivial (SURVEY.md §8(d) 'Weights'): random running stats and
    affine terms, 10% of the scales negative so max cannot be commuted through BN."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            c = m.num_features
            m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(c, generator=g) + 0.5)
            w = torch.rand(c, generator=g) + 0.5
            flip = torch.rand(c, generator=g) < 0.1
            with torch.no_grad():
                m.weight.copy_(torch.where(flip, -w, w))
                m.bias.copy_(torch.randn(c, generator=g) * 0.1)


def seeded_state_dict(model: torch.nn.Module, seed: int = 0) -> dict:
    """A reproducible, construction-order-independent set of weights for `model`: tensors are filled
    key by key in sorted key order from one torch-CPU generator (Linear ~ U(+-1/sqrt(fan_in)),
    BatchNorm as `randomize_bn_`, cls token ~ N(0,1)).  Used so the golden fixtures in tests/golden/
    need not carry a 31 MB state_dict: the fixture script and the tests both call this."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    out = {}
    for key in sorted(sd.keys()):
        ref = sd[key]
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            val = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            val = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            val = torch.rand(shape, generator=g) + 0.5
        elif leaf == "cls_token":
            val = torch.randn(shape, generator=g)
        elif leaf == "temprature":                             # CorrNet's learnable infoNCE temperature: keep its init
            val = ref.clone()
        elif ref.dim() == 2:                                   # Linear weight [out, in]
            bound = 1.0 / (shape[1] ** 0.5)
            val = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif leaf == "weight":                                 # BatchNorm scale, 10% negative
            w = torch.rand(shape, generator=g) + 0.5
            val = torch.where(torch.rand(shape, generator=g) < 0.1, -w, w)
        elif leaf == "bias":
            # Linear bias or BatchNorm shift: both small
            val = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        else:
            raise KeyError(f"unexpected state_dict entry {key} {shape}")
        out[key] = val.to(ref.dtype)
    return out


ARCH_KWARGS = {
    # factory kwargs exactly as the reference CLIs pass them
    # (training/train_rig.py:77-84, training/train_skin.py:83-88, evaluate/joint2rig.py:473)
    "jointnet_motion": dict(num_keyframes=5, chn_output=3, aggr_method="attn", motion_dim=32),
    "masknet_motion": dict(num_keyframes=5, chn_output=1, aggr_method="attn", motion_dim=32),
    "skinnet_motion": dict(nearest_bone=5, use_Dg=False, use_Lf=False, num_keyframes=5,
                           use_motion=True, motion_dim=32),
}
