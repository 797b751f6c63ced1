"""Surface-geodesic graph build on the GPU: `calc_surface_geodesic` / `get_geo_edges` of the
reference (`data_proc/common_ops.py:176-226`; offline pre-processing and `evaluate/joint2rig.py:505`) from the point
where the surface samples and their normals exist, plus a front-end that produces them.  The reference samples with
open3d (`mesh.sample_points_poisson_disk`, `estimate_normals`, :178-179), which is not in this image and whose random
stream cannot be reproduced; `sample_surface_poisson` is the stand-in: area-weighted uniform surface samples (5x
oversampled, counter-based hash RNG) thinned by farthest point sampling -- a blue-noise set like Poisson-disk sampling --
with the face normal of each sample's triangle.

    geo = calc_surface_geodesic(verts, faces)                             # sampling + geodesic matrix, like :176-208
    geo = surface_geodesic(np.asarray(samples.points), np.asarray(samples.normals), np.asarray(mesh.vertices))   # open3d samples
    geo_edge_index = get_geo_edges(geo, verts)                            # rows [i, j]; the reference's random subset rule
    geo_edge_index = geo_ball_edges(geo, radius=0.06, max_nn=15)          # deterministic variant (max_nn nearest)

numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _require_cuda_device() -> torch.device:
    """numpy inputs are processed on the current CUDA device; there is no CPU path"""
    if not torch.cuda.is_available():
        raise RuntimeError("morig_b200: a CUDA device is required (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dev):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return x.to(dev, torch.float64).contiguous()


def _device_of(x):
    if isinstance(x, np.ndarray):
        return _require_cuda_device(), True
    if not x.is_cuda:
        raise RuntimeError("morig_b200.graph_build: CUDA tensors (or numpy arrays) expected")
    return x.device, False


def surface_geodesic(pts, pts_normal, verts):
    """`calc_surface_geodesic` (common_ops.py:182-208) after the sampling: pts, pts_normal [S,3], verts [V,3] ->
    surface_geodesic [V,V] float64, bit-identical to the reference's numpy + scipy Dijkstra result."""
    lib = _lib.load()
    dev, as_numpy = _device_of(pts)
    p, nrm, v = _to_dev(pts, dev), _to_dev(pts_normal, dev), _to_dev(verts, dev)
    if p.dim() != 2 or p.shape[1] != 3 or nrm.shape != p.shape or v.dim() != 2 or v.shape[1] != 3:
        raise ValueError("surface_geodesic: pts / pts_normal must be [S,3] and verts [V,3]")
    s, nv = p.shape[0], v.shape[0]
    out = torch.empty(nv, nv, dtype=torch.float64, device=dev)
    ws_bytes = lib.morig_surface_geodesic_workspace(s, nv)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_surface_geodesic(p.data_ptr(), nrm.data_ptr(), s, v.data_ptr(), nv, out.data_ptr(),
                                              ws.data_ptr(), ws_bytes, _lib.stream_ptr()), "morig_surface_geodesic")
    return out.cpu().numpy() if as_numpy else out


def geo_ball_edges(surface_geodesic_matrix, radius=0.06, max_nn=15):
    """`get_geo_edges` (common_ops.py:214-226) after the geodesic matrix: [E,2] int64 rows (i, j), vertices in
    ascending order.  A vertex with more than `max_nn` neighbours inside the ball keeps the `max_nn` nearest (the
    reference draws a random subset there, :221)."""
    lib = _lib.load()
    dev, as_numpy = _device_of(surface_geodesic_matrix)
    g = _to_dev(surface_geodesic_matrix, dev)
    if g.dim() != 2 or g.shape[0] != g.shape[1]:
        raise ValueError("geo_ball_edges: square matrix expected")
    nv = g.shape[0]
    edges = torch.empty(nv, max_nn, 2, dtype=torch.int64, device=dev)
    deg = torch.empty(nv, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_geo_ball_edges(g.data_ptr(), nv, float(radius), int(max_nn), edges.data_ptr(),
                                            deg.data_ptr(), _lib.stream_ptr()), "morig_geo_ball_edges")
    keep = torch.arange(max_nn, device=dev).unsqueeze(0) < deg.unsqueeze(1)      # compaction of the padded rows
    out = edges[keep]
    return out.cpu().numpy() if as_numpy else out


def tpl_edges(obj_v, obj_f):
    """`get_tpl_edges(obj_v, obj_f)` (common_ops.py:15-32): [E,2] int64 rows (v, n) for every vertex v and each of its
    distinct neighbours n over the faces it belongs to.  Rows are sorted by (v, n); the reference emits a vertex's
    neighbours in python-set iteration order, the edge set is the same.  `obj_v` is only used for its type."""
    lib = _lib.load()
    as_numpy = isinstance(obj_f, np.ndarray)
    dev = _require_cuda_device() if as_numpy else obj_f.device
    if dev.type != "cuda":
        raise RuntimeError("morig_b200.graph_build: CUDA tensors (or numpy arrays) expected")
    f = (torch.from_numpy(np.ascontiguousarray(obj_f, dtype=np.int64)) if as_numpy else obj_f).to(dev, torch.int64).contiguous()
    if f.dim() != 2 or f.shape[1] != 3:
        raise ValueError("tpl_edges: faces must be [F, 3]")
    nf = f.shape[0]
    if nf == 0:
        out = torch.empty(0, 2, dtype=torch.int64, device=dev)
        return out.cpu().numpy() if as_numpy else out
    edges = torch.empty(6 * nf, 2, dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib.morig_tpl_edges_workspace(nf)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_tpl_edges(f.data_ptr(), nf, edges.data_ptr(), count.data_ptr(), ws.data_ptr(), ws_bytes,
                                       _lib.stream_ptr()), "morig_tpl_edges")
    out = edges[: int(count.item())]
    return out.cpu().numpy() if as_numpy else out


def _mesh_arrays(verts, faces, dev):
    v = _to_dev(verts, dev)
    f = (torch.from_numpy(np.ascontiguousarray(faces, dtype=np.int64)) if isinstance(faces, np.ndarray) else faces)
    return v, f.to(dev, torch.int64).contiguous()


def sample_surface_uniform(verts, faces, number_of_points, seed=0):
    """area-weighted uniform samples on the triangle mesh and the unit normals of their faces: ([S,3], [S,3]) float64"""
    lib = _lib.load()
    dev, as_numpy = _device_of(verts)
    v, f = _mesh_arrays(verts, faces, dev)
    nf, s = f.shape[0], int(number_of_points)
    pts = torch.empty(s, 3, dtype=torch.float64, device=dev)
    nrm = torch.empty(s, 3, dtype=torch.float64, device=dev)
    ws = torch.empty(2 * nf, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_sample_surface(v.data_ptr(), f.data_ptr(), nf, s, int(seed), pts.data_ptr(), nrm.data_ptr(),
                                            ws.data_ptr(), _lib.stream_ptr()), "morig_sample_surface")
    return (pts.cpu().numpy(), nrm.cpu().numpy()) if as_numpy else (pts, nrm)


def sample_surface_poisson(verts, faces, number_of_points=4000, seed=0, init_factor=5):
    """stand-in for open3d's `sample_points_poisson_disk(number_of_points)` + normals (data_proc/common_ops.py:178-179):
    init_factor x number_of_points uniform surface samples thinned to number_of_points by farthest point sampling"""
    lib = _lib.load()
    dev, as_numpy = _device_of(verts)
    v, f = _mesh_arrays(verts, faces, dev)
    s = int(number_of_points)
    raw, nrm = sample_surface_uniform(v, f, min(init_factor * s, 16384), seed)
    n_raw = raw.shape[0]
    pos32 = raw.to(torch.float32).contiguous()
    ptr = torch.tensor([0, n_raw], dtype=torch.int32, device=dev)
    out_ptr = torch.tensor([0, s], dtype=torch.int32, device=dev)
    idx = torch.empty(s, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_fps(pos32.data_ptr(), ptr.data_ptr(), out_ptr.data_ptr(), 0, 1, n_raw, idx.data_ptr(), _lib.stream_ptr()),
                   "morig_fps")
    pts, nrm = raw[idx.long()], nrm[idx.long()]
    return (pts.cpu().numpy(), nrm.cpu().numpy()) if as_numpy else (pts, nrm)


def calc_surface_geodesic(verts, faces=None, number_of_points=4000, seed=0):
    """`calc_surface_geodesic(mesh, number_of_points=4000)` -- data_proc/common_ops.py:176-208.  `verts` may be a mesh
    object with `.vertices` / `.triangles` (the reference's argument) or the vertex array with `faces` given."""
    if faces is None:
        verts, faces = np.asarray(verts.vertices), np.asarray(verts.triangles)
    pts, nrm = sample_surface_poisson(verts, faces, number_of_points, seed)
    return surface_geodesic(pts, nrm, verts)


def get_geo_edges(surface_geodesic_matrix, remesh_obj_v=None, radius=0.06, max_nn=15, seed=None):
    """`get_geo_edges` -- data_proc/common_ops.py:214-226 with the reference's subset rule: a vertex with more than max_nn
    ball members keeps `np.random.choice(members, max_nn, replace=False)`, drawn from numpy's global generator in vertex
    order exactly like the reference (seed it the same way -- or pass `seed` -- to reproduce a dataset's `_geo_e.txt`).
    Ball membership comes from the GPU; only the overflowing vertices' rows visit the host for the draw."""
    lib = _lib.load()
    dev, as_numpy = _device_of(surface_geodesic_matrix)
    g = _to_dev(surface_geodesic_matrix, dev)
    nv = g.shape[0]
    edges = torch.empty(nv, max_nn, 2, dtype=torch.int64, device=dev)
    deg = torch.empty(nv, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_geo_ball_edges(g.data_ptr(), nv, float(radius), int(max_nn), edges.data_ptr(), deg.data_ptr(),
                                            _lib.stream_ptr()), "morig_geo_ball_edges")
    deg_h = deg.cpu().numpy()
    edges_h = edges.cpu().numpy()
    full = np.nonzero(deg_h == max_nn)[0]                                        # candidates for "more than max_nn"
    rows = g[torch.from_numpy(full).to(dev)].cpu().numpy() if len(full) else np.zeros((0, nv))
    if seed is not None:
        np.random.seed(seed)
    out, k = [], 0
    for i in range(nv):
        if k < len(full) and full[k] == i:
            row = rows[k].copy()
            row[i] += 10.0                                                       # :218 no self edge
            ball = np.argwhere(row <= radius).squeeze(1)
            if len(ball) > max_nn:
                ball = np.random.choice(ball, max_nn, replace=False)             # :221
            out.append(np.stack([np.full(len(ball), i, dtype=np.int64), ball.astype(np.int64)], axis=1))
            k += 1
        else:
            out.append(edges_h[i, : deg_h[i]])
    res = np.concatenate(out, axis=0) if out else np.zeros((0, 2), dtype=np.int64)
    return res if as_numpy else torch.from_numpy(res).to(dev)
