"""Per-kernel microbenchmark on one GPU: tcgen05 3xTF32 engine vs the fp32 CUDA-core engine on the layer
shapes of the 4 x 4096-vertex jointnet_motion forward.  Prints one line per shape."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morig_b200 import engine, packing, synth  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def dense_case(M, K, N):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5
    b = torch.randn(N, generator=g).to(DEV)
    simt = packing.DenseLayer(W=packing._pack_wt(W).to(DEV), K=K, N=N, bias=b, relu=True)
    tc = packing.DenseLayer(W=simt.W, K=K, N=N, bias=b, relu=True).with_tc(W)
    tc.Wtc = tc.Wtc.to(DEV)
    C1 = torch.empty(M, N, device=DEV)
    C2 = torch.empty(M, N, device=DEV)
    t_simt = timeit(lambda: engine.dense(simt, A, 0, K, M, C=C1, ldc=N))
    t_tc = timeit(lambda: engine.dense(tc, A, 0, K, M, C=C2, ldc=N))
    ref = torch.relu(A[:2048].double() @ W.to(DEV).t() + b.double())
    fl = 2.0 * M * N * K
    print(f"dense M={M} K={K} N={N}: simt {t_simt:.3f} ms {fl / t_simt / 1e9:.1f} TF/s | tc {t_tc:.3f} ms "
          f"{fl / t_tc / 1e9:.1f} TF/s | err simt {float((C1[:2048] - ref).abs().max()):.2e} "
          f"tc {float((C2[:2048] - ref).abs().max()):.2e}", flush=True)


def edge_case(H, frames):
    data = synth.make_batch(4, 4096, seed=0).to(DEV)
    n = data.pos.shape[0]
    g = engine.graph_prep(data.geo_edge_index, n)
    gen = torch.Generator().manual_seed(2)
    pq = torch.randn(n * frames, 2 * H, generator=gen).to(DEV)
    W1 = torch.randn(H, H, generator=gen, dtype=torch.float64) / H ** 0.5
    vec = lambda: torch.randn(H, generator=gen).to(DEV)
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=vec(), scale=vec(), shift=vec(), H=H)
    br_tc = packing.EdgeBranch(W1=br.W1, b1=br.b1, scale=br.scale, shift=br.shift, H=H,
                               W1tc=packing.pack_tc_blob(W1, H, H).to(DEV))
    o1 = torch.empty(n * frames, H, device=DEV)
    o2 = torch.empty(n * frames, H, device=DEV)

    def run(b, o):
        engine.fill(o, float("-inf"))
        engine.edgeconv(b, pq, 2 * H, 0, H, g, frames, o, H, 0)
    t1 = timeit(lambda: run(br, o1))
    t2 = timeit(lambda: run(br_tc, o2))
    fl = 2.0 * g.e_max * H * H * frames
    print(f"edge H={H} frames={frames} E={g.e_max}: simt {t1:.3f} ms {fl / t1 / 1e9:.1f} TF/s | tc {t2:.3f} ms "
          f"{fl / t2 / 1e9:.1f} TF/s | max|simt-tc| {float((o1 - o2).abs().max()):.2e}", flush=True)


if __name__ == "__main__":
    R = 81920
    for M, K, N in [(R, 64, 512), (R, 288, 256), (R, 256, 1024), (R, 544, 512), (R, 832, 1024), (R, 840, 1024),
                    (R, 1024, 256), (R, 96, 64), (16384, 64, 512), (16384, 512, 64)]:
        dense_case(M, K, N)
    for H, fr in [(64, 1), (128, 5), (256, 5), (256, 1)]:
        edge_case(H, fr)
