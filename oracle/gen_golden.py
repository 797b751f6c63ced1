"""TEST INFRASTRUCTURE — regenerates tests/golden/*.npz. Runs ONLY in the build container.

It executes the reference's own, unmodified `models/rignet.py` (imported from /root/reference under
`oracle/pyg_shim.py`) on seeded synthetic batches and stores inputs + outputs.  Weights are not
stored: they come from `morig_b200.synth.seeded_state_dict(model, seed)` which tests re-run.

    python -m oracle.gen_golden          # rewrites tests/golden/
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from morig_b200 import synth  # noqa: E402
from oracle import pyg_shim  # noqa: E402

CASES = [
    # name, arch, graphs, vertices per graph, data seed, weight seed
    ("jointnet_b2_n256", "jointnet_motion", 2, 256, 10, 1),
    ("masknet_b2_n256", "masknet_motion", 2, 256, 20, 2),
    ("skinnet_b2_n256", "skinnet_motion", 2, 256, 30, 3),
    ("jointnet_b1_n1024", "jointnet_motion", 1, 1024, 0, 1),      # BASELINE.json configs[0]
    ("jointnet_mean_b3_n144", "jointnet_motion", 3, 144, 40, 4),  # aggr_method="mean" (rignet.py:92-93)
    ("masknet_max_b1_n400", "masknet_motion", 1, 400, 50, 5),     # aggr_method="max"  (rignet.py:94-95)
]


def main() -> None:
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    models = pyg_shim.import_reference_models()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    # state_dict key names + shapes of the unmodified reference (checkpoint compatibility contract)
    import json
    keys = {a: {k: list(v.shape) for k, v in models.__dict__[a](**kw).state_dict().items()}
            for a, kw in synth.ARCH_KWARGS.items()}
    with open(os.path.join(out_dir, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f)
    for name, arch, b, n, dseed, wseed in CASES:
        kw = dict(synth.ARCH_KWARGS[arch])
        if "_mean_" in name:
            kw["aggr_method"] = "mean"
        if "_max_" in name:
            kw["aggr_method"] = "max"
        model = models.__dict__[arch](**kw).eval()
        model.load_state_dict(synth.seeded_state_dict(model, wseed))
        data = synth.make_batch(b, n, seed=dseed, with_skin=(arch == "skinnet_motion"))
        with torch.no_grad():
            motion_all, motion_aggr, pred = model(data, data.pred_flow)
        blob = dict(arch=arch, graphs=b, n_vtx=n, data_seed=dseed, weight_seed=wseed,
                    aggr_method=kw.get("aggr_method", "attn"),
                    pos=data.pos.numpy(), tpl_edge_index=data.tpl_edge_index.numpy().astype(np.int32),
                    geo_edge_index=data.geo_edge_index.numpy().astype(np.int32),
                    batch=data.batch.numpy().astype(np.int32), pred_flow=data.pred_flow.numpy(),
                    motion_all=motion_all.numpy(), motion_aggr=motion_aggr.numpy(), pred=pred.numpy())
        if hasattr(data, "skin_input"):
            blob["skin_input"] = data.skin_input.numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
