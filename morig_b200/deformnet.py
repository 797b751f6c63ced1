"""`DeformNet` / `GCNDeform` / `deformnet(**kwargs)` of the reference (models/deformnet.py:13-103) on this package's kernels:
the flow producer upstream of the rigging networks (`pred_flow`).  Same constructor arguments,
`forward(data)` signature, return tuple and state_dict keys (including the reference's `mlp_tramsform`).  Inference only.

  CorrNet features -> cosine kNN (k = num_interp) of every vertex among the observed points -> similarity-weighted
  average of the point offsets (visible vertices), re-interpolated from visible to invisible vertices -> GCNDeform,
  a GCNRig with 128 / 256 / 512 channels on [flow_init | visibility].
The boolean-mask bookkeeping between the kernels (visible / invisible splits, data-dependent sizes) is index plumbing done
with torch indexing, as in the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib, engine, packing
from .basic_modules import MLP, FusedModule, GCUMotion
from .corrnet import CorrNet
from .pointnet2 import knn
from .rignet import GCNRig

__all__ = ["deformnet"]


class GCNDeform(GCNRig):
    """`GCNDeform(chn_in, chn_output, aggr='max')` -- models/deformnet.py:13-31: GCNRig's structure with wider GCUs.
    forward(pos, feature, geo_edge_index, tpl_edge_index, batch) (note the argument order of the reference)"""

    def __init__(self, chn_in, chn_output, aggr="max"):
        FusedModule.__init__(self)
        self.gcu_1 = GCUMotion(in_channels=chn_in, out_channels=128, aggr=aggr)
        self.gcu_2 = GCUMotion(in_channels=128, out_channels=256, aggr=aggr)
        self.gcu_3 = GCUMotion(in_channels=256, out_channels=512, aggr=aggr)
        self.mlp_glb = MLP([(128 + 256 + 512), 1024])
        self.mlp_tramsform = nn.Sequential(MLP([1024 + 3 + chn_in + 128 + 256 + 512, 1024, 256]), nn.Linear(256, chn_output))

    def pack(self) -> packing.GCNRigPack:
        sd = {"p." + k.replace("mlp_tramsform", "mlp_transform"): v for k, v in self._device_state().items()}
        return packing.pack_gcn_rig(sd, "p")

    def forward(self, pos, feature, geo_edge_index, tpl_edge_index, batch):
        if self.training:
            raise NotImplementedError("morig_b200.GCNDeform: inference only")
        return GCNRig.forward(self, pos, feature, tpl_edge_index, geo_edge_index, batch)


class DeformNet(FusedModule):
    def __init__(self, tau_nce, num_interp):
        super().__init__()
        self.corr_extractor = CorrNet(3, 64, temprature=tau_nce)
        self.completing = GCNDeform(chn_in=4, chn_output=3)
        self.num_interp = num_interp

    def forward(self, data):
        self._guard(data.vtx, data.pts)
        if self.training:
            raise NotImplementedError("morig_b200.DeformNet: inference only (the training path covers the rigging networks)")
        k = self.num_interp
        vtx_feature, pts_feature, pred_vismask, tau = self.corr_extractor(data, train_vismask=True)
        vb, pb = data.vtx_batch, data.pts_batch
        pred_vismask = torch.sigmoid(pred_vismask)
        for i in range(len(torch.unique(vb))):                                    # models/deformnet.py:44-46
            sel = vb == i
            v = pred_vismask[sel]
            pred_vismask[sel] = (v - v.min()) / (v.max() - v.min())
        # visible part: models/deformnet.py:49-54
        assign, sim = knn(pts_feature, vtx_feature, k, pb, vb, cosine=True, return_score=True)
        n = vtx_feature.shape[0]
        euclid = (data.pts[assign[1]] - data.vtx[assign[0]]).view(n, k, 3)
        feature_sim = torch.sum(pts_feature[assign[1]] * vtx_feature[assign[0]], dim=-1, keepdim=True).view(n, k, 1)
        feature_sim = feature_sim * pred_vismask.view(n, 1, 1)
        flow_init = (euclid * feature_sim).sum(1) / feature_sim.sum(1)
        # invisible part: models/deformnet.py:57-95
        vis_vids = (pred_vismask >= 0.5).squeeze(dim=1)
        invis_vids = (pred_vismask < 0.5).squeeze(dim=1)
        vis_f, invis_f = vtx_feature[vis_vids].contiguous(), vtx_feature[invis_vids].contiguous()
        vis_b, invis_b = vb[vis_vids], vb[invis_vids]
        vis_flow = flow_init[vis_vids]
        if invis_f.shape[0] > 0:
            assign2 = knn(vis_f, invis_f, k, vis_b, invis_b, cosine=True)
            m = invis_f.shape[0]
            sim2 = torch.sum(vis_f[assign2[1]] * invis_f[assign2[0]], dim=-1, keepdim=True).view(m, k, 1)
            invis_flow = (vis_flow[assign2[1]].view(m, k, 3) * sim2).sum(1) / sim2.sum(1)
            flow_init[invis_vids] = invis_flow
        l1_points = torch.cat((flow_init, pred_vismask), dim=-1).contiguous()
        pred_flow = self.completing(data.vtx, l1_points, data.geo_edge_index, data.tpl_edge_index, vb)
        return pred_flow, vtx_feature, pts_feature, pred_vismask, tau


def deformnet(**kwargs):
    """factory, kwargs as models/deformnet.py:101-103"""
    return DeformNet(tau_nce=kwargs["tau_nce"], num_interp=kwargs["num_interp"])
