"""Launch list of ONE graph-replayed forward at a small batch (BASELINE.json configs[0], 1 x 1024 vertices, or 1 x 4096):
run under  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file X
usage: python scripts/small_batch_profile.py [n_vtx]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morig_b200  # noqa: E402
from morig_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
kw = synth.ARCH_KWARGS["jointnet_motion"]
model = morig_b200.jointnet_motion(**kw).eval()
model.load_state_dict(synth.seeded_state_dict(model, 1))
model = model.to(dev)
data = synth.make_batch(1, n, seed=0).to(dev)
with torch.no_grad():
    for _ in range(5):
        model(data, data.pred_flow)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        model(data, data.pred_flow)
    e.record()
    torch.cuda.synchronize()
    print(f"1 x {n}: {s.elapsed_time(e) / 20:.4f} ms/forward (graph replay)")
    torch.cuda.profiler.start()
    model(data, data.pred_flow)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
