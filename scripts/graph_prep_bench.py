"""Times morig_graph_prep (CUDA events, median of 50) on the bench graphs and on a 64 K in-degree hub."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from morig_b200 import engine, synth


def time_us(ei, n, reps=50):
    ei = ei.cuda()
    for _ in range(5):
        engine.graph_prep(ei, n)
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); engine.graph_prep(ei, n); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))


out = {}
d = synth.make_batch(4, 4096, seed=0)
out["tpl 4x4096 (E=114688)"] = time_us(d.tpl_edge_index, 16384)
out["geo 4x4096 (E=262144)"] = time_us(d.geo_edge_index, 16384)
n = 70000
star = torch.stack([torch.randperm(n)[:65536], torch.full((65536,), 5)])
out["star, hub in-degree 65536"] = time_us(star, n)
print(json.dumps({"graph_prep_us": out}))
