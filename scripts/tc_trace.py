"""Role timeline of CTA 0 of one tcgen05 launch (debug): prints per-role event durations."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from morig_b200 import _lib, engine, packing, synth
import tc_microbench as t
lib = _lib.load()
lib.morig_debug_set_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(3 * 2048 * 2, dtype=torch.int64, device="cuda:0")

def run(case):
    lib.morig_debug_set_trace(None)
    case()          # warm (also prints timing)
    buf.zero_()
    lib.morig_debug_set_trace(buf.data_ptr())
    case_once[0] = True
    case()
    lib.morig_debug_set_trace(None)
    torch.cuda.synchronize()
    tr = buf.cpu()[:3 * 2048].view(3, 1024, 2)
    names = {1: "P wait-free", 2: "P got-free", 3: "P stored", 10: "E wait-acc", 11: "E got-acc", 12: "E released",
             13: "E ld-done", 14: "E sts-done", 15: "E lds-done", 16: "E math-done", 17: "E flush-done", 20: "C wait-accE", 21: "C got-accE", 22: "C got-A", 23: "C got-B", 24: "C committed", 25: "C got-prevMMA"}
    for role, nm in enumerate(["producer(w0)", "control", "epilogue(w8)"]):
        ev = [(int(a), int(b)) for a, b in tr[role].tolist() if a != 0]
        if not ev: continue
        t0 = ev[0][1]
        print(f"--- {nm}: {len(ev)} events")
        prev = t0
        line = []
        for tag, clk in ev[:120]:
            line.append(f"{names.get(tag, tag)}+{clk - prev}")
            prev = clk
        print("  ".join(line))
        # average gaps by tag
        import collections
        acc = collections.defaultdict(list); prev = ev[0][1]
        for tag, clk in ev[1:]:
            acc[tag].append(clk - prev); prev = clk
        print({names.get(k, k): (len(v), sum(v) // len(v)) for k, v in acc.items()}, "total cycles", ev[-1][1] - ev[0][1])

case_once = [False]
which = sys.argv[1] if len(sys.argv) > 1 else "edge256"
if which == "edge256": run(lambda: t.edge_case(256, 5))
elif which == "edge128": run(lambda: t.edge_case(128, 5))
else: run(lambda: t.dense_case(81920, 832, 1024))
