"""Training-mode forwards of the rigging networks: the same modules, `model.train()`, with
train-mode BatchNorm statistics (over E rows inside the edge MLPs, over N rows in the vertex MLPs), running-statistics
updates, and a differentiable graph of `autograd_ops` functions so that `loss.backward()` of
training/train_rig.py:136-195 / training/train_skin.py:139-183 runs on this package's kernels.

Structure follows the reference module by module (each function cites it).  Differences that are exact in real arithmetic:
the first Linear of every edge MLP is evaluated per vertex (`W0 [x_i, x_j - x_i] + b0 = (Wa - Wb) x_i + b0 + Wb x_j`), and
the x / pos halves of an EdgeConvMotion message are max-reduced separately and concatenated afterwards.  Unlike the
inference path the five key-frame passes are NOT batched: each `motionNet` call has its own batch statistics and moves
the running statistics once, exactly as the reference's loop does (models/rignet.py:85-88).
"""
from __future__ import annotations

import torch

from . import _lib, engine
from . import autograd_ops as A
from . import train_ops as T


def mlp_block(x, blk):
    """one `Seq(Linear, ReLU, BatchNorm1d)` of `MLP` -- models/basic_modules.py:31-36"""
    lin, bn = blk[0], blk[2]
    w = lin.weight
    if x.shape[1] > w.shape[1]:        # input came from ConcatColsPadded: zero weight columns for the zero input columns
        w = torch.cat([w, w.new_zeros(w.shape[0], x.shape[1] - w.shape[1])], dim=1)
    y = A.LinReluBN.apply(x, w, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum)
    _count_batch(bn)
    return y


def _count_batch(bn) -> None:
    if bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1


def mlp(x, seq):
    for blk in seq:
        x = mlp_block(x, blk)
    return x


def edge_branch(x, seq, g):
    """one 2-layer edge MLP of an EdgeConv on graph `g`, max-aggregated: [N, C] -> [N, H]
    (models/basic_modules.py:193-199 message + PyG max aggregation)"""
    lin0, bn0 = seq[0][0], seq[0][2]
    c = lin0.weight.shape[1] // 2
    wa, wb = lin0.weight[:, :c], lin0.weight[:, c:]
    wpq = torch.cat([wa - wb, wb], dim=0)                                  # parameter-sized algebra (autograd: torch)
    bpq = torch.cat([lin0.bias, torch.zeros_like(lin0.bias)])
    pq = A.Linear.apply(x, wpq, bpq)                                       # [N, 2H] = [P | Q]
    h0 = A.EdgeGatherRelu.apply(pq, g)                                     # [E, H]
    a0 = A.BNTrain.apply(h0, bn0.weight, bn0.bias, bn0.running_mean, bn0.running_var, bn0.momentum)
    _count_batch(bn0)
    y1 = mlp_block(a0, seq[1])                                             # [E, H]
    out, _ = A.SegMax.apply(y1, g.rowptr, g.n)
    return out


def _as_matrix(x):
    return x.unsqueeze(-1) if x.dim() == 1 else x


def edge_conv_motion_parts(mod, pos, x, g):
    """EdgeConvMotion (models/basic_modules.py:185-199) -> ([N, H], [N, Dp]) halves of its output"""
    return edge_branch(_as_matrix(x), mod.nn_x, g), edge_branch(pos, mod.nn_pos, g)


def gcu_motion(mod, pos, x, gt, gg):
    """GCUMotion.forward -- models/basic_modules.py:214-219"""
    xt, pt = edge_conv_motion_parts(mod.edge_conv_tpl, pos, x, gt)
    xg, pg = edge_conv_motion_parts(mod.edge_conv_geo, pos, x, gg)
    return mlp(A.ConcatCols.apply(xt, pt, xg, pg), mod.mlp)


def gcu(mod, x, gt, gg):
    """GCU.forward -- models/basic_modules.py:172-177"""
    x = _as_matrix(x)
    a = edge_branch(x, mod.edge_conv_tpl.nn_pos, gt)
    b = edge_branch(x, mod.edge_conv_geo.nn_pos, gg)
    return mlp(A.ConcatCols.apply(a, b), mod.mlp)


def batch_ptr(binfo: engine.BatchInfo) -> torch.Tensor:
    if getattr(binfo, "ptr", None) is None:
        binfo.ptr = T.seg_ptr(binfo.batch32, binfo.n_graphs)
    return binfo.ptr


def gcn_rig(mod, pos, feature, gt, gg, binfo):
    """GCNRig.forward -- models/rignet.py:58-67"""
    feature = _as_matrix(feature)
    x1 = gcu_motion(mod.gcu_1, pos, feature, gt, gg)
    x2 = gcu_motion(mod.gcu_2, pos, x1, gt, gg)
    x3 = gcu_motion(mod.gcu_3, pos, x2, gt, gg)
    x4 = mlp(A.ConcatCols.apply(x1, x2, x3), mod.mlp_glb)
    ptr = batch_ptr(binfo)
    xg, _ = A.SegMax.apply(x4, ptr, binfo.n_graphs)
    xgr = A.RowGather.apply(xg, binfo.batch32, ptr)
    x5 = A.ConcatColsPadded.apply(xgr, pos, feature, x1, x2, x3)
    h = mlp(x5, mod.mlp_transform[0])
    head = mod.mlp_transform[1]
    return A.Linear.apply(h, head.weight, head.bias)


def temporal_attn(mod, x):
    """TemporalAttn.forward -- models/rignet.py:36-46.  Only row 0 of the attention output feeds the rest (:45), so only
    the cls query is evaluated; its gradient reaches w_qs / cls_token exactly as in the reference."""
    n, t, c = x.shape
    hd = mod.w_ks.weight.shape[0]
    d = hd // mod.num_heads
    xf = x.reshape(n * t, c)
    kx = A.Linear.apply(xf, mod.w_ks.weight, None).view(n, t, hd)
    vx = A.Linear.apply(xf, mod.w_vs.weight, None).view(n, t, hd)
    cls = mod.cls_token.view(1, c)
    q0 = A.Linear.apply(cls, mod.w_qs.weight, None)
    kc = A.Linear.apply(cls, mod.w_ks.weight, None)
    vc = A.Linear.apply(cls, mod.w_vs.weight, None)
    r0 = A.AttnCls.apply(q0, kc, vc, kx, vx, d)
    o = A.Linear.apply(r0, mod.w_o.weight, None)
    return mlp(o, mod.feedforward)


def graph_for_training(graphs: engine.GraphCache, edge_index, n) -> engine.Graph:
    """CSR of the edge list plus its edge count on the host: the per-edge activations are materialised as [E', H]
    tensors in training, which needs E' (one 4-byte read per new edge list; the reference's remove_self_loops syncs at
    the same place)."""
    g = graphs.get(edge_index, n)
    g.join()
    if getattr(g, "e_real", None) is None:
        g.e_real = int(g.rowptr[n].item())
    return g


def motion_encode(model, pos, flow, gt, gg, binfo, dim):
    """key-frame loop -- models/rignet.py:84-89,117-122,196-201"""
    frames = []
    for t in range(model.num_keyframes):
        m = gcn_rig(model.motionNet, pos, flow[:, 3 * t:3 * t + 3], gt, gg, binfo)
        frames.append(A.Normalize.apply(m))
    return A.ConcatCols.apply(*frames).view(pos.shape[0], model.num_keyframes, dim)     # == torch.stack(frames, dim=1)


def aggregate(model, motion_all, aggr_method):
    """models/rignet.py:90-98"""
    if aggr_method == "attn":
        a = temporal_attn(model.aggragator, motion_all)
    elif aggr_method == "mean":
        a = A.FrameMean.apply(motion_all)
    elif aggr_method == "max":
        a = A.FrameMax.apply(motion_all)
    else:
        raise NotImplementedError(aggr_method)
    return A.Normalize.apply(a)


def _train_inputs(model, data, input_flow):
    pos = _lib.require_cuda(data.pos, "data.pos")
    flow = _lib.require_cuda(input_flow, "input_flow")
    n = pos.shape[0]
    gt = graph_for_training(model._graphs, data.tpl_edge_index, n)
    gg = graph_for_training(model._graphs, data.geo_edge_index, n)
    binfo = model._batches.get(data.batch, data)
    return pos, flow, gt, gg, binfo


def joint_mask_forward(model, data, input_flow):
    """JointNetMotion / MaskNetMotion.forward in train mode -- models/rignet.py:82-100, 115-133"""
    pos, flow, gt, gg, binfo = _train_inputs(model, data, input_flow)
    motion_all = motion_encode(model, pos, flow, gt, gg, binfo, 32)
    motion_aggr = aggregate(model, motion_all, model.aggr_method)
    pred = gcn_rig(getattr(model, model._head_name), pos, motion_aggr, gt, gg, binfo)
    return motion_all, motion_aggr, pred


def skin_inner(mod, data, pos, motion, gt, gg, binfo):
    """SkinNet_inner.forward -- models/rignet.py:158-182"""
    from . import packing
    skin = _lib.require_cuda(data.skin_input, "data.skin_input")
    cols = packing.skin_columns(skin.shape[1], mod.num_nearest_bone, mod.use_Dg, mod.use_Lf)
    n = pos.shape[0]
    raw = torch.empty(n, 3 + len(cols), dtype=pos.dtype, device=pos.device)
    cols_t = torch.tensor(cols, dtype=torch.int32, device=pos.device)
    engine.gather_cols(pos, 3, 0, 0, None, 3, n, 1, raw, raw.shape[1], 0)
    engine.gather_cols(skin, skin.shape[1], 0, 0, cols_t, len(cols), n, 1, raw, raw.shape[1], 3)
    x1 = gcu_motion(mod.gcu1, raw, motion, gt, gg)
    ptr = batch_ptr(binfo)
    xg, _ = A.SegMax.apply(mlp(x1, mod.multi_layer_tranform2), ptr, binfo.n_graphs)
    x2 = gcu_motion(mod.gcu2, raw, x1, gt, gg)
    x3 = gcu_motion(mod.gcu3, raw, x2, gt, gg)
    xgr = A.RowGather.apply(xg, binfo.batch32, ptr)
    h = mlp(A.ConcatCols.apply(x3, xgr), mod.cls_branch[0])
    head = mod.cls_branch[1]
    return A.Linear.apply(h, head.weight, head.bias)


def skin_forward(model, data, input_flow):
    """SkinMotion.forward in train mode -- models/rignet.py:194-205"""
    pos, flow, gt, gg, binfo = _train_inputs(model, data, input_flow)
    motion_all = motion_encode(model, pos, flow, gt, gg, binfo, model.motion_dim)
    motion_aggr = aggregate(model, motion_all, "attn")
    pred = skin_inner(model.skinNet, data, pos, motion_aggr, gt, gg, binfo)
    return motion_all, motion_aggr, pred
