"""Joint-extraction post-process, GPU side: drop-in for `utils.cluster_utils.meanshift_cluster`
of the reference (utils/cluster_utils.py:14-35), which `evaluate/eval_rigging.py:91` runs on the shifted vertices
right after the jointnet / masknet forward.  The reference materialises the N x N kernel matrix in fp64 numpy every
iteration; `morig_meanshift_step` streams it.  Same loop, same stopping rule, fp64 like the reference's pipeline
(float32 inputs are promoted).  `nms_meanshift` (:38-63) follows: the ball statistics and a bit matrix of the balls are
computed in parallel, the greedy sweep runs in one CTA (csrc/postproc.cu); `flip` (utils/mst_utils.py:294-313) is the
index bookkeeping that eval_rigging.py:95 applies to the surviving modes.
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


def meanshift_cluster(pts_in, bandwidth, weights=None, max_iter=20, return_iters=False):
    """`meanshift_cluster(pts_in, bandwidth, weights=None, max_iter=20)` — utils/cluster_utils.py:14.
    pts_in [N,3] and weights [N] / [N,1] as numpy arrays (returns numpy, like the reference) or CUDA tensors
    (returns a CUDA float64 tensor).  There is no CPU path: a CUDA device is required."""
    lib = _lib.load()
    as_numpy = isinstance(pts_in, np.ndarray)
    dev = _require_cuda_device() if as_numpy else pts_in.device
    if dev.type != "cuda":
        raise RuntimeError("morig_b200.cluster_utils.meanshift_cluster: CUDA tensors (or numpy arrays) expected")
    pts = (torch.from_numpy(np.ascontiguousarray(pts_in, dtype=np.float64)) if as_numpy else pts_in).to(dev, torch.float64)
    pts = pts.contiguous().clone()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError(f"pts_in must be [N, 3], got {tuple(pts.shape)}")
    n = pts.shape[0]
    w = None
    if weights is not None:
        w = (torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)) if isinstance(weights, np.ndarray)
             else weights).to(dev, torch.float64).reshape(-1).contiguous()
        if w.shape[0] != n:
            raise ValueError(f"weights must have {n} entries, got {w.shape[0]}")
    nxt = torch.empty_like(pts)
    d2 = torch.empty(n, dtype=torch.float64, device=dev)
    diff_sq = torch.zeros(1, dtype=torch.float64, device=dev)
    diff, num_iter = 1e10, 1                                   # loop of utils/cluster_utils.py:22-34
    with torch.cuda.device(dev):
        while diff > 1e-3 and num_iter < max_iter:
            _lib.check(lib.morig_meanshift_step(pts.data_ptr(), _lib.ptr(w), float(bandwidth), n, nxt.data_ptr(),
                                                d2.data_ptr(), diff_sq.data_ptr(), _lib.stream_ptr()),
                       "morig_meanshift_step")
            diff = float(diff_sq.item()) ** 0.5               # the reference reads diff on the host every iteration too
            pts, nxt = nxt, pts
            num_iter += 1
    out = pts.cpu().numpy() if as_numpy else pts
    return (out, num_iter - 1) if return_iters else out


def _as_cuda_f64(x, dev):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)) if isinstance(x, np.ndarray) else x
    return t.to(dev, torch.float64).contiguous()


def nms_meanshift(pts_in, attn, bandwidth, thrd_density, thrd_attn=0.7):
    """`nms_meanshift(pts_in, attn, bandwidth, thrd_density, thrd_attn=0.7)` -- utils/cluster_utils.py:38-63: the modes
    that survive non-maximum suppression, in input order (numpy in -> numpy out, CUDA tensors in -> CUDA tensor out).
    Points are visited by decreasing neighbour count; the reference breaks ties with numpy's default (unstable) argsort,
    i.e. leaves them unspecified -- here equal counts go from the higher index down (np.argsort(kind='stable')[::-1])."""
    lib = _lib.load()
    as_numpy = isinstance(pts_in, np.ndarray)
    dev = _require_cuda_device() if as_numpy else pts_in.device
    if dev.type != "cuda":
        raise RuntimeError("morig_b200.cluster_utils.nms_meanshift: CUDA tensors (or numpy arrays) expected")
    pts = _as_cuda_f64(pts_in, dev)
    n = pts.shape[0]
    if n == 0:
        return pts_in
    a = _as_cuda_f64(attn, dev).reshape(-1)
    if a.shape[0] != n:
        raise ValueError(f"attn must have {n} entries, got {a.shape[0]}")
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    nbytes = lib.morig_nms_meanshift_workspace(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.morig_nms_meanshift(pts.data_ptr(), a.data_ptr(), n, float(bandwidth), float(thrd_density),
                                           float(thrd_attn), keep.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()),
                   "morig_nms_meanshift")
    if as_numpy:
        return pts_in[keep.cpu().numpy().astype(bool)]
    return pts_in[keep.bool()]                                 # data-dependent size: the one sync, as in the reference


def flip(pred_joints):
    """`flip(pred_joints)` -- utils/mst_utils.py:294-313: symmetrise the predicted joints by reflecting the left half
    space (x < -0.02) to the right and snapping the middle band (|x| <= 0.02) onto the plane.  Returns
    (joints [L + M + L, 3], side indicator).  Index bookkeeping on a handful of joints: done where the data lives."""
    if isinstance(pred_joints, np.ndarray):
        left = pred_joints[pred_joints[:, 0] < -2e-2]
        mid = pred_joints[np.abs(pred_joints[:, 0]) <= 2e-2].copy()
        mid[:, 0] = 0.0
        right = left.copy()
        right[:, 0] = -right[:, 0]
        side = np.concatenate((-np.ones(len(left)), np.zeros(len(mid)), np.ones(len(right))), axis=0)
        return np.concatenate((left, mid, right), axis=0), side
    left = pred_joints[pred_joints[:, 0] < -2e-2]
    mid = pred_joints[pred_joints[:, 0].abs() <= 2e-2].clone()
    mid[:, 0] = 0.0
    right = left.clone()
    right[:, 0] = -right[:, 0]
    side = torch.cat((-torch.ones(len(left)), torch.zeros(len(mid)), torch.ones(len(right)))).to(pred_joints.device)
    return torch.cat((left, mid, right), dim=0), side
