"""BASELINE.json configs 2-4 (parity-test cases, not bench lines): device-resident forward time of the three networks
at their configured sizes, CUDA events, graph-replayed steps with an L2 flush between them.  One JSON line per config."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morig_b200  # noqa: E402
from morig_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for arch, b, n in (("jointnet_motion", 4, 4096), ("masknet_motion", 8, 4096), ("skinnet_motion", 4, 8192)):
    kw = synth.ARCH_KWARGS[arch]
    model = getattr(morig_b200, arch)(**kw).eval()
    model.load_state_dict(synth.seeded_state_dict(model, 1))
    model = model.to(dev)
    data = synth.make_batch(b, n, seed=0, with_skin=(arch == "skinnet_motion")).to(dev)
    with torch.no_grad():
        for _ in range(4):
            model(data, data.pred_flow)
        torch.cuda.synchronize()
        ms = 0.0
        steps = 10
        for _ in range(steps):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            model(data, data.pred_flow)
            e.record()
            torch.cuda.synchronize()
            ms += s.elapsed_time(e)
    print(json.dumps({"config": f"{arch} forward, batch={b} x {n}-vertex meshes, 1xB200", "ms_per_step": ms / steps,
                      "meshes_per_s": b * steps / (ms * 1e-3)}), flush=True)
    del model, data
    torch.cuda.empty_cache()
