"""Max |cuda - oracle| of the three networks on a GPU (prints one line per network)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from morig_b200 import synth  # noqa: E402

for arch, b, n in [("jointnet_motion", 2, 1024), ("masknet_motion", 1, 2048), ("skinnet_motion", 2, 1024)]:
    kw = synth.ARCH_KWARGS[arch]
    errs = []
    for wseed in (11, 12):
        data = synth.make_batch(b, n, seed=100 + wseed, with_skin=(arch == "skinnet_motion"))
        model = helpers.build_model(arch, kw, wseed, "cuda:0")
        expect = helpers.oracle_forward(arch, kw, model, data, data.pred_flow)
        with torch.no_grad():
            out = model(data.to("cuda:0"), data.pred_flow.to("cuda:0"))
        errs.append([helpers.max_abs_diff(o, e) for o, e in zip(out, expect)])
    print(arch, " ".join(f"{max(e[i] for e in errs):.2e}" for i in range(3)), "(motion_all motion_aggr pred; tol 1e-4)",
          flush=True)
