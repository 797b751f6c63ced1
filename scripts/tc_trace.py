"""Role timeline of CTA 0 of one tcgen05 launch (debug): prints per-role event gaps.
usage: python scripts/tc_trace.py {edge256|edge128|dense} {f16|tf32} [frames]"""
import collections
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morig_b200 import _lib, engine, packing, synth

DEV = "cuda:0"
lib = _lib.load()
lib.morig_debug_set_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(3 * 2048 * 2, dtype=torch.int64, device=DEV)
NAMES = {1: "P wait-free", 2: "P got-free", 3: "P stored", 10: "E wait-acc", 11: "E got-acc", 12: "E released",
         13: "E ld-done", 14: "E sts-done", 16: "E walk-done", 17: "E tile-done", 20: "C wait-accE", 21: "C got-accE",
         22: "C got-A", 23: "C got-B", 24: "C committed", 25: "C got-prevMMA"}
WS = engine.Workspace()


def make_case(which, kind, frames):
    gen = torch.Generator().manual_seed(2)
    if which.startswith("edge"):
        H = int(which[4:])
        data = synth.make_batch(4, 4096, seed=0).to(DEV)
        n = data.pos.shape[0]
        g = engine.graph_prep(data.geo_edge_index, n)
        pq = torch.randn(n * frames, 2 * H, generator=gen).to(DEV)
        W1 = torch.randn(H, H, generator=gen, dtype=torch.float64) / H ** 0.5
        vec = lambda: torch.randn(H, generator=gen).to(DEV)
        scale = vec()
        blob, w_inv = packing.pack_edge_tc_blob(W1, scale.cpu(), kind)
        br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=vec(), scale=scale, shift=vec(), H=H,
                                W1tc=blob.to(DEV), tc_kind=kind, tc_w_inv=w_inv)
        o = torch.empty(n * frames, H, device=DEV)
        return lambda: engine.edgeconv(br, pq, 2 * H, 0, H, g, frames, o, H, 0)
    M, K, N = 81920, 832, 1024
    A = torch.randn(M, K, generator=gen).to(DEV)
    W = torch.randn(N, K, generator=gen, dtype=torch.float64) / K ** 0.5
    layer = packing.DenseLayer(W=packing._pack_wt(W).to(DEV), K=K, N=N, relu=True).with_tc(W, kind)
    layer.Wtc = layer.Wtc.to(DEV)
    C = torch.empty(M, N, device=DEV)
    return lambda: engine.dense(layer, A, 0, K, M, C=C, ldc=N)


which = sys.argv[1] if len(sys.argv) > 1 else "edge256"
kind = packing.KIND_F16 if (len(sys.argv) < 3 or sys.argv[2] == "f16") else packing.KIND_TF32
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 5
case = make_case(which, kind, frames)
with engine.forward_scope(WS, DEV):
    for _ in range(3):
        case()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); case(); e.record(); torch.cuda.synchronize()
    print(f"{which} kind={'f16' if kind else 'tf32'} frames={frames}: {s.elapsed_time(e):.3f} ms")
    lib.morig_debug_set_trace(buf.data_ptr())
    case()
    lib.morig_debug_set_trace(None)
torch.cuda.synchronize()
tr = buf.cpu().view(6, 1024, 2)[:3]          # roles are 2048 int64 slots (= 1024 events) apart
for role, nm in enumerate(["producer(w0)", "control", "epilogue(w4)"]):
    ev = [(int(a), int(b)) for a, b in tr[role].tolist() if a != 0]
    if not ev:
        continue
    print(f"--- {nm}: {len(ev)} events")
    prev = ev[0][1]
    line = []
    for tag, clk in ev[40:100]:
        line.append(f"{NAMES.get(tag, tag)}+{clk - prev}")
        prev = clk
    print("  ".join(line[1:]))
    acc = collections.defaultdict(list)
    prev = ev[0][1]
    for tag, clk in ev[1:]:
        acc[tag].append(clk - prev)
        prev = clk
    print({NAMES.get(k, k): (len(v), sum(v) // len(v)) for k, v in acc.items()}, "total cycles", ev[-1][1] - ev[0][1])
