"""Joint-extraction post-process, GPU side: drop-in for `utils.cluster_utils.meanshift_cluster`
of the reference (utils/cluster_utils.py:14-35), which `evaluate/eval_rigging.py:91` runs on the shifted vertices
right after the jointnet / masknet forward.  The reference materialises the N x N kernel matrix in fp64 numpy every
iteration; `morig_meanshift_step` streams it.  Same loop, same stopping rule, fp64 like the reference's pipeline
(float32 inputs are promoted).  The non-maximum suppression that follows (`nms_meanshift`, :38-63) is sequential,
data dependent and tie-broken by numpy's unstable argsort; it is not rebuilt here.
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
