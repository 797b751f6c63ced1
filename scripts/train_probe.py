"""Debug probe: loss trajectory of a few SGD steps and the distance between the tensor-core and the fp32 training GEMMs
(whole-network gradient), 2 x 256-vertex jointnet.  usage: python scripts/train_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers  # noqa: E402
from morig_b200 import synth  # noqa: E402

DEV = "cuda:0"


def run(tc: str, lr: float, steps: int):
    os.environ["MORIG_TRAIN_TC"] = tc
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 3, DEV).train()
    data = synth.make_batch(2, 256, seed=5).to(DEV)
    g = torch.Generator().manual_seed(9)
    target = torch.tanh(torch.randn(512, 3, generator=g)).to(DEV) * 0.1
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    losses, grads = [], None
    for i in range(steps):
        opt.zero_grad()
        _, _, pred = model(data, data.pred_flow)
        loss = (torch.tanh(pred) - target).pow(2).mean()
        loss.backward()
        if i == 0:
            grads = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).clone()
        opt.step()
        losses.append(round(float(loss.detach()), 5))
    return losses, grads


for lr in (1e-5, 1e-6, 1e-7):
    l1, g1 = run("1", lr, 8)
    l0, g0 = run("0", lr, 8)
    print(f"lr={lr} tc  ", l1)
    print(f"lr={lr} fp32", l0)
print("grad rel L2 (tc vs fp32):", float((g1 - g0).norm() / g0.norm()), "norm", float(g0.norm()))
l1b, g1b = run("1", 1e-3, 2)
print("tc run-to-run grad rel L2:", float((g1 - g1b).norm() / g1.norm()))
