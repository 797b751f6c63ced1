"""`CorrNet` / `corrnet(**kwargs)` of the reference (models/corrnet.py:10-77) on this package's kernels: the per-vertex
and per-point feature extractor whose cosine correspondences feed DeformNet.  Same constructor
arguments, `forward(data, train_vismask, random_start=True)` signature, return tuple and state_dict keys.  Inference only.

  vertex branch   four GCU layers (the C_p = 0 case of the fused EdgeConv kernels) + pooled global feature + MLP head
  point branch    PointNet++ set abstraction / feature propagation (pointnet2.py: fps, ball query, PointConv, knn_interpolate)
  visibility      1-nearest point feature of every vertex by cosine similarity (morig_knn_topk) -> MLP
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from . import train_ops as T
from .basic_modules import GCU, MLP, FusedModule
from .pointnet2 import DenseStack, FPModule, GlobalSAModule, SAModule, knn

__all__ = ["corrnet"]


class CorrNet(FusedModule):
    def __init__(self, input_feature, output_feature, temprature, aggr="max"):
        super().__init__()
        self.input_feature = input_feature
        self.output_feature = output_feature
        self.temprature = nn.Parameter(torch.Tensor([temprature]))
        self.vtx_gcu_1 = GCU(in_channels=3, out_channels=32, aggr=aggr)
        self.vtx_gcu_2 = GCU(in_channels=32, out_channels=64, aggr=aggr)
        self.vtx_gcu_3 = GCU(in_channels=64, out_channels=256, aggr=aggr)
        self.vtx_gcu_4 = GCU(in_channels=256, out_channels=512, aggr=aggr)
        self.vtx_mlp_glb = MLP([(32 + 64 + 256 + 512), 1024])
        self.vtx_mlp = nn.Sequential(MLP([1024 + 3 + 32 + 64 + 256 + 512, 1024, 256]), nn.Linear(256, output_feature))
        self.pts_sa1_module = SAModule(0.5, 0.12, MLP([input_feature, 32, 32, 64]), max_num_neighbors=64)
        self.pts_sa2_module = SAModule(0.25, 0.25, MLP([64 + 3, 64, 64, 128]), max_num_neighbors=64)
        self.pts_sa3_module = SAModule(0.25, 0.5, MLP([128 + 3, 256, 256, 256]), max_num_neighbors=64)
        self.pts_sa4_module = GlobalSAModule(MLP([256 + 3, 256, 256, 512]))
        self.pts_fp4_module = FPModule(1, MLP([512 + 256, 256, 256]))
        self.pts_fp3_module = FPModule(3, MLP([256 + 128, 256, 128]))
        self.pts_fp2_module = FPModule(3, MLP([128 + 64, 128, 64]))
        self.pts_fp1_module = FPModule(3, MLP([64, 64, 64]))
        self.pts_mlp = nn.Sequential(MLP([64, 64]), nn.Linear(64, output_feature))
        self.lin_vismask = nn.Sequential(MLP([2 * output_feature + 1, 256, 128, 64]), nn.Linear(64, 1))

    def forward(self, data, train_vismask, random_start=True):
        self._guard(data.vtx, data.pts)
        if self.training:
            raise NotImplementedError("morig_b200.CorrNet: inference only (the training path covers the rigging networks)")
        vtx = _lib.require_cuda(data.vtx, "data.vtx")
        pts = _lib.require_cuda(data.pts, "data.pts")
        geo, tpl, vb, pb = data.geo_edge_index, data.tpl_edge_index, data.vtx_batch, data.pts_batch
        # ---- vertex branch: models/corrnet.py:40-49
        x_1 = self.vtx_gcu_1(vtx, tpl, geo)
        x_2 = self.vtx_gcu_2(x_1, tpl, geo)
        x_3 = self.vtx_gcu_3(x_2, tpl, geo)
        x_4 = self.vtx_gcu_4(x_3, tpl, geo)
        x_5 = DenseStack(self, "vtx_mlp_glb", self.vtx_mlp_glb)(T.concat_cols([x_1, x_2, x_3, x_4]))
        vb32 = vb.to(torch.int32)
        n_graphs = getattr(data, "num_graphs", None) or int(vb[-1].item()) + 1
        ptr = T.seg_ptr(vb32, n_graphs)
        x_global, _ = T.segmax_fwd(x_5, ptr, n_graphs)                           # scatter_max(x_5, vtx_batch)
        x_6 = T.concat_cols([T.row_gather(x_global, vb32), vtx, x_1, x_2, x_3, x_4])
        out_vtx = T.normalize_fwd(DenseStack(self, "vtx_mlp", self.vtx_mlp)(x_6))
        # ---- point branch: models/corrnet.py:51-62
        sa0 = (None, pts, pb)
        sa1 = self.pts_sa1_module(*sa0, random_start)
        sa2 = self.pts_sa2_module(*sa1, random_start)
        sa3 = self.pts_sa3_module(*sa2, random_start)
        sa4 = self.pts_sa4_module(*sa3)
        fp4 = self.pts_fp4_module(*sa4, *sa3)
        fp3 = self.pts_fp3_module(*fp4, *sa2)
        fp2 = self.pts_fp2_module(*fp3, *sa1)
        out_pts, _, _ = self.pts_fp1_module(*fp2, *sa0)
        out_pts = T.normalize_fwd(DenseStack(self, "pts_mlp", self.pts_mlp)(out_pts))
        # ---- visibility mask: models/corrnet.py:64-75 (CUDA branch: cosine 1-NN of every vertex among its sample's points)
        if train_vismask:
            assign, sim = knn(out_pts, out_vtx, 1, pb, vb, cosine=True, return_score=True)
            out_combine = T.concat_cols([out_vtx, T.row_gather(out_pts, assign[1].to(torch.int32)), sim])
            out_vismask = DenseStack(self, "lin_vismask", self.lin_vismask)(out_combine)
        else:
            out_vismask = None
        return out_vtx, out_pts, out_vismask, self.temprature


def corrnet(**kwargs):
    """factory, kwargs as models/corrnet.py:80-82"""
    return CorrNet(input_feature=kwargs["input_feature"], output_feature=kwargs["output_feature"], temprature=kwargs["temprature"])
