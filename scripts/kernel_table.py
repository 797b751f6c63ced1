"""Per-kernel table (CUDA events around every launch, plain stream launches) of one network at one size.
usage: python scripts/kernel_table.py {jointnet_motion|masknet_motion|skinnet_motion} B N"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morig_b200  # noqa: E402
from morig_b200 import engine, synth  # noqa: E402

arch, b, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
kw = synth.ARCH_KWARGS[arch]
model = getattr(morig_b200, arch)(**kw).eval()
model.load_state_dict(synth.seeded_state_dict(model, 1))
model = model.to(dev)
data = synth.make_batch(b, n, seed=0, with_skin=(arch == "skinnet_motion")).to(dev)
counter, prof = engine.LaunchCounter(), engine.KernelTimer()
with torch.no_grad():
    model(data, data.pred_flow)
    engine.set_hooks(counter, prof)
    model(data, data.pred_flow)
    counter.reset(); prof.reset()
    for _ in range(5):
        model(data, data.pred_flow)
ks = prof.summary()
tot = sum(k["total_ms"] for k in ks)
print(f"{arch} {b} x {n}: {tot / 5:.3f} ms of kernels per step, {counter.count // 5} launches")
for k in ks[:28]:
    print(f"{k['kernel']:52s} n={k['launches'] // 5:3d} avg={k['avg_ms']:.4f} share={k['total_ms'] / tot * 100:5.1f}% tf={k['tflops']:.1f} gbs={k['gbs']:.0f}")
