"""One training step (4 x 4096-vertex jointnet_motion, forward + backward) between cudaProfilerStart/Stop after a warm step:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python scripts/train_profile.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import morig_b200
from morig_b200 import synth

dev = torch.device("cuda:0")
kw = synth.ARCH_KWARGS["jointnet_motion"]
model = morig_b200.jointnet_motion(**kw)
model.load_state_dict(synth.seeded_state_dict(model, 1))
model = model.to(dev).train()
data = synth.make_batch(4, 4096, seed=0).to(dev)


def step():
    model.zero_grad(set_to_none=True)
    _, _, pred = model(data, data.pred_flow)
    torch.tanh(pred).pow(2).mean().backward()


step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
print("wall ms / step:", 1e3 * (time.perf_counter() - t0))
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
