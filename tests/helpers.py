"""Shared helpers of the test-suite (fixture loading, model construction)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

import morig_b200
from morig_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT_KEYS = ("motion_all", "motion_aggr", "pred")
# north_star tolerance: "per-vertex offsets and attention within 1e-4 fp32" (absolute)
TOL = 1e-4


def golden_names():
    """network fixtures (jointnet_* / masknet_* / skinnet_*); other fixtures are loaded by the tests that use them"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*net_*.npz")))


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    arch = str(z["arch"])
    kw = dict(synth.ARCH_KWARGS[arch])
    if "aggr_method" in kw:
        kw["aggr_method"] = str(z["aggr_method"])
    data = synth.Batch(pos=torch.from_numpy(z["pos"]),
                       tpl_edge_index=torch.from_numpy(z["tpl_edge_index"]).long(),
                       geo_edge_index=torch.from_numpy(z["geo_edge_index"]).long(),
                       batch=torch.from_numpy(z["batch"]).long(),
                       pred_flow=torch.from_numpy(z["pred_flow"]))
    if "skin_input" in z:
        data.skin_input = torch.from_numpy(z["skin_input"])
    expect = tuple(torch.from_numpy(z[k]) for k in OUT_KEYS)
    return arch, kw, int(z["weight_seed"]), data, expect


def build_model(arch: str, kw: dict, weight_seed: int, device="cpu"):
    model = getattr(morig_b200, arch)(**kw).eval()
    model.load_state_dict(synth.seeded_state_dict(model, weight_seed))
    return model.to(device)


def oracle_forward(arch: str, kw: dict, model, data, flow):
    from oracle import rignet_port
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    args = dict(num_keyframes=kw["num_keyframes"])
    if arch == "skinnet_motion":
        args.update(nearest_bone=kw["nearest_bone"], use_Dg=kw["use_Dg"], use_Lf=kw["use_Lf"])
    else:
        args.update(aggr_method=kw["aggr_method"])
    with torch.no_grad():
        return rignet_port.FORWARDS[arch](sd, data, flow, **args)


def max_abs_diff(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())
