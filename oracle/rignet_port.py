"""TEST INFRASTRUCTURE — the CPU oracle ("port") for the hot path. Never imported by `morig_b200`.

A self-contained restatement, in plain torch-CPU tensor ops driven by a reference-keyed
`state_dict`, of the forward of the reference's three motion-aware rigging networks. It exists
because `/root/reference` cannot travel to the GPU box: tests, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` check (and time) against this file.

Pinned (in the build container) by `tests/test_oracle_pinning.py`, which runs the reference's own
unmodified `models/rignet.py` under `oracle/pyg_shim.py` on the same inputs and demands
bit-identical outputs, and by the committed fixtures in `tests/golden/` that were produced by that
unmodified code (`oracle/gen_golden.py`).  The reference repo itself holds no tests or golden
vectors for this path, so that is the strongest pin available.

The op sequence deliberately mirrors the reference step for step (per-edge gathers, explicit
concatenations, Linear -> ReLU -> BatchNorm, scatter-max) so that its CPU timing is
representative of the reference's CPU path.  Eval-mode by default (running BatchNorm statistics); under
`with training_mode():` it is the training oracle (batch statistics, running-statistics updates, autograd).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# Training oracle: with TRAINING set (see `training_mode`) BatchNorm uses batch statistics and moves
# the running statistics stored in `sd` in place, exactly like the reference's modules in `.train()` mode; gradients
# come from torch autograd on the `sd` tensors that require grad.
TRAINING = False


class training_mode:
    def __enter__(self):
        global TRAINING
        self.prev, TRAINING = TRAINING, True
        return self

    def __exit__(self, *exc):
        global TRAINING
        TRAINING = self.prev
        return False


def _mlp(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """`MLP(channels)` = stack of (Linear, ReLU, BatchNorm1d) — models/basic_modules.py:31-36.
    Keys `{prefix}.{layer}.0.{weight,bias}` (Linear) and `{prefix}.{layer}.2.*` (BatchNorm)."""
    layer = 0
    while f"{prefix}.{layer}.0.weight" in sd:
        p = f"{prefix}.{layer}"
        x = F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
        x = F.relu(x)
        x = F.batch_norm(x, sd[p + ".2.running_mean"], sd[p + ".2.running_var"],
                         sd[p + ".2.weight"], sd[p + ".2.bias"], training=TRAINING, momentum=0.1, eps=1e-5)
        layer += 1
    return x


def _segment_max(msg: torch.Tensor, dst: torch.Tensor, n: int) -> torch.Tensor:
    """PyG `aggr='max'` -> torch_scatter max over the target index; empty segment -> 0
    (models/basic_modules.py:180-181,190; third-party semantics restated in oracle/pyg_shim.py)."""
    lowest = torch.finfo(msg.dtype).min
    idx2 = dst.unsqueeze(1).expand_as(msg)
    with torch.no_grad():
        out = torch.full((n, msg.shape[1]), lowest, dtype=msg.dtype)
        out = out.scatter_reduce(0, idx2, msg, reduce="amax", include_self=True)
        seen = torch.zeros(n, dtype=torch.bool)
        seen[dst] = True
        out = torch.where(seen.unsqueeze(1), out, torch.zeros_like(out))
    if torch.is_grad_enabled() and msg.requires_grad and msg.shape[0] > 0:
        # gradient rule of torch_scatter's max: only the FIRST maximal row of a segment receives it
        rows = torch.arange(msg.shape[0]).unsqueeze(1).expand_as(msg)
        cand = torch.where(msg.detach() == out[dst], rows, torch.full_like(rows, msg.shape[0]))
        arg = torch.full((n, msg.shape[1]), msg.shape[0], dtype=torch.long)
        arg = arg.scatter_reduce(0, idx2, cand, reduce="amin", include_self=True)
        picked = msg.gather(0, arg.clamp(max=msg.shape[0] - 1))
        out = torch.where(arg < msg.shape[0], picked, torch.zeros_like(picked))
    return out


def normalized_edges(edge_index: torch.Tensor, n: int) -> torch.Tensor:
    """strip self loops, then append one loop per vertex — models/basic_modules.py:188-189."""
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(n, dtype=edge_index.dtype)
    return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)


def edge_conv_motion(sd, prefix: str, pos: torch.Tensor, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """`EdgeConvMotion.forward/message/update` — models/basic_modules.py:185-199.
    Row 0 of edge_index is the source j, row 1 the target i (flow source_to_target)."""
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    n = x.shape[0]
    ei = normalized_edges(edge_index, n)
    j, i = ei[0], ei[1]
    x_i, x_j = x.index_select(0, i), x.index_select(0, j)
    p_i, p_j = pos.index_select(0, i), pos.index_select(0, j)
    feat_x = _mlp(sd, prefix + ".nn_x", torch.cat([x_i, x_j - x_i], dim=1))
    feat_p = _mlp(sd, prefix + ".nn_pos", torch.cat([p_i, p_j - p_i], dim=1))
    return _segment_max(torch.cat([feat_x, feat_p], dim=1), i, n)


def gcu_motion(sd, prefix: str, pos, x, tpl_ei, geo_ei) -> torch.Tensor:
    """`GCUMotion.forward` — models/basic_modules.py:214-219."""
    a = edge_conv_motion(sd, prefix + ".edge_conv_tpl", pos, x, tpl_ei)
    b = edge_conv_motion(sd, prefix + ".edge_conv_geo", pos, x, geo_ei)
    return _mlp(sd, prefix + ".mlp", torch.cat([a, b], dim=1))


def edge_conv(sd, prefix: str, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """`EdgeConv.forward/message/update` — models/basic_modules.py:148-159 (single MLP `nn_pos` applied to x)."""
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    n = x.shape[0]
    ei = normalized_edges(edge_index, n)
    j, i = ei[0], ei[1]
    x_i, x_j = x.index_select(0, i), x.index_select(0, j)
    return _segment_max(_mlp(sd, prefix + ".nn_pos", torch.cat([x_i, x_j - x_i], dim=1)), i, n)


def gcu(sd, prefix: str, pos, tpl_ei, geo_ei) -> torch.Tensor:
    """`GCU.forward` — models/basic_modules.py:172-177."""
    a = edge_conv(sd, prefix + ".edge_conv_tpl", pos, tpl_ei)
    b = edge_conv(sd, prefix + ".edge_conv_geo", pos, geo_ei)
    return _mlp(sd, prefix + ".mlp", torch.cat([a, b], dim=1))


def graph_max(x: torch.Tensor, batch: torch.Tensor) -> torch.Tensor:
    """`scatter_max(x, batch, dim=0)[0]` — models/rignet.py:63,176."""
    nb = int(batch.max()) + 1
    return _segment_max(x, batch, nb)


def gcn_rig(sd, prefix: str, pos, feature, tpl_ei, geo_ei, batch) -> torch.Tensor:
    """`GCNRig.forward` — models/rignet.py:58-67."""
    x1 = gcu_motion(sd, prefix + ".gcu_1", pos, feature, tpl_ei, geo_ei)
    x2 = gcu_motion(sd, prefix + ".gcu_2", pos, x1, tpl_ei, geo_ei)
    x3 = gcu_motion(sd, prefix + ".gcu_3", pos, x2, tpl_ei, geo_ei)
    x4 = _mlp(sd, prefix + ".mlp_glb", torch.cat([x1, x2, x3], dim=1))
    xg = graph_max(x4, batch)
    xg = torch.repeat_interleave(xg, torch.bincount(batch), dim=0)
    x5 = torch.cat([xg, pos, feature, x1, x2, x3], dim=1)
    h = _mlp(sd, prefix + ".mlp_transform.0", x5)
    return F.linear(h, sd[prefix + ".mlp_transform.1.weight"], sd[prefix + ".mlp_transform.1.bias"])


def temporal_attn(sd, prefix: str, x: torch.Tensor, num_heads: int = 2) -> torch.Tensor:
    """`TemporalAttn.forward` — models/rignet.py:36-46 (with transpose_qkv/transpose_output :22-34).
    x: [N, T, C]; a learnable cls token is prepended and only its output row is kept."""
    n = x.shape[0]
    tok = torch.cat([sd[prefix + ".cls_token"].expand(n, -1, -1), x], dim=1)        # [N, T+1, C]
    q = F.linear(tok, sd[prefix + ".w_qs.weight"])
    k = F.linear(tok, sd[prefix + ".w_ks.weight"])
    v = F.linear(tok, sd[prefix + ".w_vs.weight"])

    def split(t):   # [N, L, heads*d] -> [N*heads, L, d]
        t = t.reshape(t.shape[0], t.shape[1], num_heads, -1).permute(0, 2, 1, 3)
        return t.reshape(-1, t.shape[2], t.shape[3])

    q, k, v = split(q), split(k), split(v)
    att = torch.bmm(q, k.transpose(-2, -1))
    att = F.softmax(att / math.sqrt(k.size(-1)), dim=-1)
    r = torch.bmm(att, v)                                                            # [N*heads, L, d]
    r = r.reshape(-1, num_heads, r.shape[1], r.shape[2]).permute(0, 2, 1, 3)
    r = r.reshape(r.shape[0], r.shape[1], -1)                                        # [N, L, heads*d]
    r = F.linear(r, sd[prefix + ".w_o.weight"])
    return _mlp(sd, prefix + ".feedforward", r[:, 0, :])


def motion_encoder(sd, data, input_flow, num_keyframes: int):
    """key-frame loop shared by the three nets — models/rignet.py:84-89,117-122,196-201."""
    frames = []
    for t in range(num_keyframes):
        m = gcn_rig(sd, "motionNet", data.pos, input_flow[:, 3 * t:3 * t + 3],
                    data.tpl_edge_index, data.geo_edge_index, data.batch)
        frames.append(F.normalize(m, dim=1))
    return torch.stack(frames, dim=1)


def _aggregate(sd, motion_all, aggr_method: str):
    """models/rignet.py:90-98."""
    if aggr_method == "attn":
        a = temporal_attn(sd, "aggragator", motion_all)
    elif aggr_method == "mean":
        a = motion_all.mean(dim=1)
    elif aggr_method == "max":
        a = motion_all.max(dim=1)[0]
    else:
        raise NotImplementedError(aggr_method)
    return F.normalize(a, dim=1)


def jointnet_motion_forward(sd, data, input_flow, num_keyframes: int = 5, aggr_method: str = "attn"):
    """`JointNetMotion.forward` — models/rignet.py:82-100. Returns (motion_all, motion_aggr, pred_shift)."""
    motion_all = motion_encoder(sd, data, input_flow, num_keyframes)
    motion_aggr = _aggregate(sd, motion_all, aggr_method)
    pred = gcn_rig(sd, "jointnet", data.pos, motion_aggr, data.tpl_edge_index, data.geo_edge_index, data.batch)
    return motion_all, motion_aggr, pred


def masknet_motion_forward(sd, data, input_flow, num_keyframes: int = 5, aggr_method: str = "attn"):
    """`MaskNetMotion.forward` — models/rignet.py:115-133. Returns (motion_all, motion_aggr, pred_mask)."""
    motion_all = motion_encoder(sd, data, input_flow, num_keyframes)
    motion_aggr = _aggregate(sd, motion_all, aggr_method)
    pred = gcn_rig(sd, "masknet", data.pos, motion_aggr, data.tpl_edge_index, data.geo_edge_index, data.batch)
    return motion_all, motion_aggr, pred


def skin_columns(width: int, nearest_bone: int, use_Dg: bool, use_Lf: bool) -> torch.Tensor:
    """Which columns of `data.skin_input` survive the selection of models/rignet.py:159-171."""
    cols = torch.arange(width)
    if use_Dg and use_Lf:
        return cols[: 8 * nearest_bone]
    if use_Dg and not use_Lf:
        return cols[cols % 8 != 7][: 7 * nearest_bone]
    if use_Lf and not use_Dg:
        return cols[cols % 8 != 6][: 7 * nearest_bone]
    cols = cols[cols % 8 != 7]
    cols = cols[torch.arange(cols.numel()) % 7 != 6]
    return cols[: 6 * nearest_bone]


def skinnet_inner(sd, prefix: str, data, motion, nearest_bone: int, use_Dg: bool, use_Lf: bool):
    """`SkinNet_inner.forward` — models/rignet.py:158-182."""
    samples = data.skin_input[:, skin_columns(data.skin_input.shape[1], nearest_bone, use_Dg, use_Lf)]
    raw = torch.cat([data.pos, samples], dim=1)
    tpl, geo = data.tpl_edge_index, data.geo_edge_index
    x1 = gcu_motion(sd, prefix + ".gcu1", raw, motion, tpl, geo)
    xg = graph_max(_mlp(sd, prefix + ".multi_layer_tranform2", x1), data.batch)
    x2 = gcu_motion(sd, prefix + ".gcu2", raw, x1, tpl, geo)
    x3 = gcu_motion(sd, prefix + ".gcu3", raw, x2, tpl, geo)
    xg = torch.repeat_interleave(xg, torch.bincount(data.batch), dim=0)
    h = _mlp(sd, prefix + ".cls_branch.0", torch.cat([x3, xg], dim=1))
    return F.linear(h, sd[prefix + ".cls_branch.1.weight"], sd[prefix + ".cls_branch.1.bias"])


def skinnet_motion_forward(sd, data, input_flow, num_keyframes: int = 5, nearest_bone: int = 5,
                           use_Dg: bool = False, use_Lf: bool = False):
    """`SkinMotion.forward` — models/rignet.py:194-205. Returns (motion_all, motion_aggr, skin_cls_pred)."""
    motion_all = motion_encoder(sd, data, input_flow, num_keyframes)
    motion_aggr = _aggregate(sd, motion_all, "attn")
    pred = skinnet_inner(sd, "skinNet", data, motion_aggr, nearest_bone, use_Dg, use_Lf)
    return motion_all, motion_aggr, pred


FORWARDS = {"jointnet_motion": jointnet_motion_forward,
            "masknet_motion": masknet_motion_forward,
            "skinnet_motion": skinnet_motion_forward}
