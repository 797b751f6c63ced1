"""Measurement of the mean-shift row: GPU time per iteration with CUDA events, pair
evaluations/s and fp64 FLOP/s of `meanshift_step_kernel`, next to the reference numpy path (oracle port, which is
bit-identical to utils/cluster_utils.py:14-35) on the host cores.  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morig_b200 import _lib  # noqa: E402
from oracle import cluster_port  # noqa: E402

N_HALF = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
c = rng.uniform(-0.4, 0.4, size=(20, 3))
pts = c[rng.integers(0, 20, N_HALF)] + rng.normal(0, 0.02, size=(N_HALF, 3))
pts = np.concatenate([pts, pts * np.array([[-1, 1, 1]])], axis=0)
w = np.tile(rng.uniform(0.05, 1.0, size=(N_HALF, 1)).astype(np.float32), (2, 1))
n = pts.shape[0]
dev = torch.device("cuda:0")
lib = _lib.load()
p = torch.from_numpy(pts).to(dev)
q = torch.empty_like(p)
wd = torch.from_numpy(w.astype(np.float64).reshape(-1)).to(dev)
d2 = torch.empty(n, dtype=torch.float64, device=dev)
dsq = torch.zeros(1, dtype=torch.float64, device=dev)


def step():
    _lib.check(lib.morig_meanshift_step(p.data_ptr(), wd.data_ptr(), 0.05, n, q.data_ptr(), d2.data_ptr(), dsq.data_ptr(),
                                        _lib.stream_ptr()), "morig_meanshift_step")


for _ in range(3):
    step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
s.record()
for _ in range(iters):
    step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / iters
flops_per_pair = 17.0           # 3 sub, 3 mul/fma for y, 1 sub + max + mul for k, 4 fma accumulations (2 flops each)
t0 = time.perf_counter()
_, ref_iters = cluster_port.meanshift_cluster(pts[:4096], 0.05, w[:4096], 4, return_iters=True)
cpu_s_per_iter = (time.perf_counter() - t0) / ref_iters
print(json.dumps({"row": "meanshift_cluster (utils/cluster_utils.py:14-35)", "n_points": n, "gpu_ms_per_iteration": ms,
                  "pair_evals_per_s": n * n / (ms * 1e-3), "fp64_tflops": n * n * flops_per_pair / (ms * 1e-3) / 1e12,
                  "fp64_nominal_peak_tflops": 37.2, "cpu_ref_s_per_iteration_at_4096_points": cpu_s_per_iter,
                  "cpu_ref_scaled_to_n_s": cpu_s_per_iter * (n / 4096.0) ** 2, "cpu_cores": os.cpu_count()}))
