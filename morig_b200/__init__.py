"""morig_b200 — B200-native (sm_100a) forward path of MoRig's motion-aware rigging networks.

    import morig_b200
    model = morig_b200.jointnet_motion(num_keyframes=5, chn_output=3, aggr_method="attn").cuda().eval()
    model.load_state_dict(reference_checkpoint["state_dict"])        # same keys as the reference
    motion_all, motion_aggr, pred = model(data, data.pred_flow)      # same call as training/train_rig.py:223

    pipe = morig_b200.HostPipeline(model)                            # host batches in, pinned host results out,
    for outs in pipe.run(host_batches): ...                          # copies overlapped with the forwards

or, to swap the networks inside an unmodified reference checkout:

    import models, morig_b200
    morig_b200.install(models)            # models.__dict__['jointnet_motion'] etc. now build the CUDA modules
"""
from __future__ import annotations

from .basic_modules import GCU, MLP, EdgeConv, EdgeConvMotion, GCUMotion
from .corrnet import CorrNet, corrnet
from .deformnet import DeformNet, GCNDeform, deformnet
from .pipeline import HostPipeline
from .pointnet2 import FPModule, GlobalSAModule, SAModule
from .rignet import (GCNRig, JointNetMotion, MaskNetMotion, SkinMotion, SkinNet_inner, TemporalAttn,
                     jointnet_motion, masknet_motion, skinnet_motion)

__all__ = ["jointnet_motion", "masknet_motion", "skinnet_motion", "JointNetMotion", "MaskNetMotion", "SkinMotion",
           "SkinNet_inner", "GCNRig", "TemporalAttn", "GCUMotion", "EdgeConvMotion", "GCU", "EdgeConv", "MLP", "install",
           "HostPipeline", "corrnet", "deformnet", "CorrNet", "DeformNet", "GCNDeform", "SAModule", "GlobalSAModule", "FPModule"]

__version__ = "0.1.0"


def install(models_module, flow_producer: bool = False) -> None:
    """Replace the rigging-network entries of the reference's `models` registry
    (`models/__init__.py:1-3`; looked up as `models.__dict__[args.arch]` at training/train_rig.py:83).
    `flow_producer=True` also replaces `corrnet` / `deformnet` (inference-only here; their training stays on the
    reference's modules)."""
    for name in ("jointnet_motion", "masknet_motion", "skinnet_motion", "JointNetMotion", "MaskNetMotion",
                 "SkinMotion", "SkinNet_inner", "GCNRig", "TemporalAttn"):
        setattr(models_module, name, globals()[name])
    if flow_producer:
        for name in ("corrnet", "deformnet", "CorrNet", "DeformNet", "GCNDeform"):
            setattr(models_module, name, globals()[name])
    sub = getattr(models_module, "rignet", None)
    if sub is not None:
        for name in ("jointnet_motion", "masknet_motion", "skinnet_motion", "JointNetMotion", "MaskNetMotion",
                     "SkinMotion"):
            setattr(sub, name, globals()[name])
