"""ncu target: one launch of each dominant kernel shape of the 4 x 4096-vertex jointnet_motion forward (fp16-split
operand kind) between cudaProfilerStart/Stop, after two warm launches.  Run under
    ncu --set full --import-source on --profile-from-start off -k regex:"gemm_kernel|edge_mma" ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morig_b200 import engine, packing, synth  # noqa: E402

DEV = "cuda:0"
WS = engine.Workspace()
KIND = packing.KIND_F16


def profiled(fn):
    with engine.forward_scope(WS, DEV):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


def dense_case(M, K, N):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5
    b = torch.randn(N, generator=g).to(DEV)
    layer = packing.DenseLayer(W=packing._pack_wt(W).to(DEV), K=K, N=N, bias=b, relu=True).with_tc(W, KIND)
    layer.Wtc = layer.Wtc.to(DEV)
    C = torch.empty(M, N, device=DEV)
    profiled(lambda: engine.dense(layer, A, 0, K, M, C=C, ldc=N))


def edge_case(H, frames):
    data = synth.make_batch(4, 4096, seed=0).to(DEV)
    n = data.pos.shape[0]
    g = engine.graph_prep(data.geo_edge_index, n)
    gen = torch.Generator().manual_seed(2)
    pq = torch.randn(n * frames, 2 * H, generator=gen).to(DEV)
    W1 = torch.randn(H, H, generator=gen, dtype=torch.float64) / H ** 0.5
    vec = lambda: torch.randn(H, generator=gen).to(DEV)
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=vec(), scale=vec(), shift=vec(), H=H)
    if H >= 64:
        blob, w_inv = packing.pack_edge_tc_blob(W1, br.scale.cpu(), KIND)
        br.W1tc, br.tc_kind, br.tc_w_inv = blob.to(DEV), KIND, w_inv
    o = torch.full((n * frames, H), float("-inf"), device=DEV)
    profiled(lambda: engine.edgeconv(br, pq, 2 * H, 0, H, g, frames, o, H, 0))


if __name__ == "__main__":
    which = sys.argv[1:] or ["e256", "e128", "d840", "e32"]
    for w in which:
        if w == "e256":
            edge_case(256, 5)
        elif w == "e128":
            edge_case(128, 5)
        elif w == "e32":
            edge_case(32, 5)
        elif w == "e16":
            edge_case(16, 1)
        elif w == "d840":
            dense_case(81920, 840, 1024)
        elif w == "d1024":
            dense_case(81920, 1024, 256)
