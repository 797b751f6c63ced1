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
    ref = torch.relu(A[:2048].double() @ W.to(DEV).t() + b.double())
    fl = 2.0 * M * N * K
    line = f"dense M={M} K={K} N={N}:"
    for name, kind in (("simt", None), ("tf32", packing.KIND_TF32), ("f16", packing.KIND_F16)):
        layer = simt
        if kind is not None:
            layer = packing.DenseLayer(W=simt.W, K=K, N=N, bias=b, relu=True).with_tc(W, kind)
            layer.Wtc = layer.Wtc.to(DEV)
        C = torch.empty(M, N, device=DEV)
        with engine.forward_scope(WS, DEV):          # operand range computed once, outside the timed launches
            engine.dense(layer, A, 0, K, M, C=C, ldc=N)
            t = timeit(lambda: engine.dense(layer, A, 0, K, M, C=C, ldc=N))
        line += f" {name} {t:.3f} ms {fl / t / 1e9:.1f} TF/s err {float((C[:2048] - ref).abs().max()):.2e} |"
    print(line, flush=True)


def edge_case(H, frames):
    data = synth.make_batch(4, 4096, seed=0).to(DEV)
    n = data.pos.shape[0]
    g = engine.graph_prep(data.geo_edge_index, n)
    gen = torch.Generator().manual_seed(2)
    pq = torch.randn(n * frames, 2 * H, generator=gen).to(DEV)
    W1 = torch.randn(H, H, generator=gen, dtype=torch.float64) / H ** 0.5
    vec = lambda: torch.randn(H, generator=gen).to(DEV)
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=vec(), scale=vec(), shift=vec(), H=H)
    fl = 2.0 * g.e_max * H * H * frames
    line = f"edge H={H} frames={frames} E={g.e_max}:"
    o_ref = None
    for name, kind in (("simt", None), ("tf32", packing.KIND_TF32), ("f16", packing.KIND_F16)):
        b = br
        if kind is not None:
            blob, w_inv = packing.pack_edge_tc_blob(W1, br.scale.cpu(), kind)
            b = packing.EdgeBranch(W1=br.W1, b1=br.b1, scale=br.scale, shift=br.shift, H=H, W1tc=blob.to(DEV),
                                   tc_kind=kind, tc_w_inv=w_inv)
        o = torch.empty(n * frames, H, device=DEV)

        def run():
            engine.fill(o, float("-inf"))
            engine.edgeconv(b, pq, 2 * H, 0, H, g, frames, o, H, 0)
        with engine.forward_scope(WS, DEV):
            run()
            t = timeit(run)
        o_ref = o if o_ref is None else o_ref
        line += f" {name} {t:.3f} ms {fl / t / 1e9:.1f} TF/s d {float((o - o_ref).abs().max()):.2e} |"
    print(line, flush=True)


WS = engine.Workspace()

if __name__ == "__main__":
    R = 81920
    only_edge = len(sys.argv) > 1 and sys.argv[1] == "edge"
    for M, K, N in [] if only_edge else [(R, 64, 512), (R, 288, 256), (R, 256, 1024), (R, 544, 512), (R, 832, 1024), (R, 840, 1024),
                    (R, 1024, 256), (R, 96, 64), (16384, 64, 512), (16384, 512, 64)]:
        dense_case(M, K, N)
    for H, fr in [(64, 1), (128, 5), (256, 5), (256, 1)]:
        edge_case(H, fr)
