"""TEST INFRASTRUCTURE -- regenerates tests/golden/deformnet_*.npz.  Runs ONLY in the build container: it executes the
reference's own, unmodified `models/deformnet.py` + `models/corrnet.py` (PointNet++ modules of models/basic_modules.py)
on the CPU under the third-party stand-ins of oracle/pyg_shim.py + oracle/pointops_port.py and stores inputs + outputs.
Weights come from `morig_b200.synth.seeded_state_dict`; the random FPS start points come from the seeded torch generator.

    python -m oracle.gen_golden_deform
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from morig_b200 import synth  # noqa: E402
from oracle import pointops_port, pyg_shim  # noqa: E402

CASES = [("deformnet_b2_v400_p300", 2, 400, 300, 11, 5, 123)]      # name, graphs, vertices, points, data seed, weight seed, rng seed


def reference_models():
    pointops_port.install()
    models = pyg_shim.import_reference_models()
    pointops_port.install()                                      # patch the already imported reference modules too
    return models


def run_reference(models, data, wseed, rng_seed):
    net = models.deformnet(tau_nce=0.07, num_interp=5).eval()
    net.load_state_dict(synth.seeded_state_dict(net, wseed))
    torch.manual_seed(rng_seed)
    # The reference switches on torch.cuda.is_available(): on a GPU it calls torch_cluster's batched `radius` / `knn`
    # (models/basic_modules.py:78-80, models/corrnet.py:65-67), on the CPU its own `radius_cpu`, which ignores `batch`
    # and samples randomly when a ball holds more than max_num_neighbors points.  The GPU branch is the one this package
    # replaces, so the reference is made to take it here (the stand-ins of oracle/pointops_port.py run on CPU tensors).
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        with torch.no_grad():
            pred_flow, vtx_f, pts_f, vis, _ = net(data)
    finally:
        torch.cuda.is_available = real
    return pred_flow, vtx_f, pts_f, vis


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    models = reference_models()
    for name, b, nv, npts, dseed, wseed, rseed in CASES:
        data = synth.make_deform_batch(b, nv, npts, seed=dseed)
        pred_flow, vtx_f, pts_f, vis = run_reference(models, data, wseed, rseed)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, graphs=b, n_vtx=nv, n_pts=npts, data_seed=dseed, weight_seed=wseed, rng_seed=rseed,
                            vtx=data.vtx.numpy(), pts=data.pts.numpy(), vtx_batch=data.vtx_batch.numpy().astype(np.int32),
                            pts_batch=data.pts_batch.numpy().astype(np.int32),
                            tpl_edge_index=data.tpl_edge_index.numpy().astype(np.int32),
                            geo_edge_index=data.geo_edge_index.numpy().astype(np.int32),
                            pred_flow=pred_flow.numpy(), vtx_feature=vtx_f.numpy(), pts_feature=pts_f.numpy(), pred_vismask=vis.numpy())
        print(f"{name}: wrote {os.path.getsize(path) / 1e3:.0f} kB; |pred_flow| max {float(pred_flow.abs().max()):.3f}, "
              f"visible {int((vis >= 0.5).sum())}/{vis.numel()}")


if __name__ == "__main__":
    main()
