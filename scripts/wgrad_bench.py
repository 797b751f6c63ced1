"""fp32 weight-gradient kernel (csrc/train.cu wgrad_kernel) on the training shapes of the 4 x 4096-vertex jointnet step"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morig_b200 import train_ops as T

out = {}
for M, N, K in ((409600, 256, 256), (409600, 128, 128), (16384, 1024, 1864), (16384, 1024, 832), (409600, 32, 32), (16384, 512, 544)):
    dy, x = torch.randn(M, N, device="cuda"), torch.randn(M, K, device="cuda")
    for _ in range(3):
        T.wgrad(dy, x, True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        T.wgrad(dy, x, True)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    out[f"M={M} N={N} K={K}"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 2)}
print(json.dumps(out))
