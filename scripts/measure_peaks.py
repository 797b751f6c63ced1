"""Compute peaks the driver does not measure (SURVEY.md §8(d) 'Peaks to divide by'): cuBLAS fp32 (FFMA), TF32 and fp16
8192^3 matmuls, best of 10 (burst) and back to back for 3 s (sustained), the way MEASURED_PEAKS.json measures bf16."""
import json
import time

import torch

dev = torch.device("cuda:0")
n = 8192
out = {}
for name, dtype, tf32 in (("fp32_ffma", torch.float32, False), ("tf32", torch.float32, True), ("fp16", torch.float16, False),
                          ("bf16", torch.bfloat16, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a @ b; e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, iters = time.perf_counter(), 0
    s.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(10):
            a @ b
        iters += 10
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    out[name] = {"burst_tflops": 2 * n ** 3 / (best * 1e-3) / 1e12, "sustained_tflops": 2 * n ** 3 * iters / (s.elapsed_time(e) * 1e-3) / 1e12}
print(json.dumps(out))
