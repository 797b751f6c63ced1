"""Host-side mirror of the graph-conv building blocks of the reference (`models/basic_modules.py`):
same constructor arguments, same `forward` signatures, same `state_dict` keys — but `forward` runs
the fused sm_100a kernels of `libmorig_b200.so`.  Parameters live in ordinary torch containers purely
so that reference checkpoints load unchanged (`training/train_rig.py:95`); the containers themselves
are never executed.  `model.eval()` runs the fused inference kernels (no autograd); `model.train()` runs the training
path (train-mode BatchNorm, autograd functions of `autograd_ops.py`, see `train_forward.py`).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import _lib, engine, packing


def MLP(channels: List[int], batch_norm: bool = True) -> nn.Sequential:
    """Parameter container with the key layout of the reference's `MLP`
    (`models/basic_modules.py:31-36`): `{i}.0` Linear, `{i}.1` ReLU, `{i}.2` BatchNorm1d(momentum=0.1)."""
    if not batch_norm:
        raise NotImplementedError("the rigging networks only use batch_norm=True")
    blocks = []
    for c_in, c_out in zip(channels[:-1], channels[1:]):
        blocks.append(nn.Sequential(nn.Linear(c_in, c_out), nn.ReLU(), nn.BatchNorm1d(c_out, momentum=0.1)))
    return nn.Sequential(*blocks)


def _check_rows(who: str, n: int, **tensors) -> None:
    for name, t in tensors.items():
        if t.dim() < 1 or t.shape[0] != n:
            raise ValueError(f"{who}: `{name}` must have {n} rows, got shape {tuple(t.shape)}")


def _check_width(who: str, **pairs) -> None:
    """the reference raises a shape error from `Linear`; raw kernels would read out of bounds instead"""
    for name, (t, want) in pairs.items():
        if t.dim() != 2 or t.shape[1] != want:
            raise ValueError(f"{who}: `{name}` must be [N, {want}], got shape {tuple(t.shape)}")


# bumped whenever any module's parameters may have changed; captured CUDA graphs remember the epoch they
# were built in and are rebuilt when it moved (they bake packed-weight pointers)
WEIGHTS_EPOCH = [0]


def _cuda_device_of(args):
    """device of the first CUDA tensor among a forward's arguments (tensors, or batch objects with `.pos`)"""
    for a in args:
        t = a if torch.is_tensor(a) else getattr(a, "pos", None)
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    return None


class FusedModule(nn.Module):
    """Common behaviour of the drop-in modules: lazily packed device weights that are dropped whenever
    the parameters may have changed (`load_state_dict`, `.to()`, `.train()`, or any tracked in-place edit of a
    parameter / buffer: optimizer steps, `nn.init`, `load_state_dict` on a plain child container), a reusable
    workspace and loud failure outside the supported regime."""

    def __init__(self):
        super().__init__()
        self._packed = None
        self._packed_key = None
        self._packed_fp = None
        self._fp_tensors = None
        self._ws = engine.Workspace()
        self._graphs = engine.GraphCache()
        self._batches = engine.BatchCache()
        self._register_load_state_dict_pre_hook(self._drop_packed_hook)

    # -- cache invalidation ----------------------------------------------------------------------
    def _drop_packed_hook(self, *args, **kwargs):
        self.invalidate_packed()

    def invalidate_packed(self):
        WEIGHTS_EPOCH[0] += 1
        for m in self.modules():
            if isinstance(m, FusedModule):
                m._packed = None
                m._packed_key = None
                m._packed_fp = None
                m._fp_tensors = None

    def __call__(self, *args, **kwargs):
        # raw kernel launches go to the current stream of the CURRENT device: make that the inputs' device, as
        # PyTorch ops do implicitly (a model on cuda:1 must work while cuda:0 is current)
        dev = _cuda_device_of(args)
        if dev is None or dev.index == torch.cuda.current_device():
            return super().__call__(*args, **kwargs)
        with torch.cuda.device(dev):
            return super().__call__(*args, **kwargs)

    def weights_fingerprint(self):
        """(storage address, in-place version) of every parameter and buffer below this module: changes with
        optimizer steps, `nn.init.*_`, `load_state_dict` (also on children), `.to()`.  Untracked writes through
        `.data` are the one thing it cannot see -- call `invalidate_packed()` after those."""
        ts = self._fp_tensors
        if ts is None:
            ts = self._fp_tensors = list(self.state_dict(keep_vars=True).values())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_packed()
        self._ws.clear()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        self.invalidate_packed()
        return super().train(mode)

    # -- helpers ---------------------------------------------------------------------------------
    def _device_state(self):
        return {k: v for k, v in self.state_dict(keep_vars=True).items()}

    def _guard(self, *tensors: torch.Tensor):
        _lib.load()
        for t in tensors:
            if not t.is_cuda:
                raise RuntimeError(f"morig_b200 runs on CUDA tensors only (got {t.device}); there is no CPU path")

    def _packed_for(self, key, builder):
        fp = self.weights_fingerprint()
        if self._packed is None or self._packed_key != key or self._packed_fp != fp:
            if self._packed is not None and self._packed_fp != fp:
                WEIGHTS_EPOCH[0] += 1          # captured CUDA graphs bake pointers of the old pack
            with torch.no_grad():
                self._packed = builder()
            self._packed_key = key
            self._packed_fp = fp
        return self._packed


class EdgeConvMotion(FusedModule):
    """`EdgeConvMotion(nn_x, nn_pos, aggr='max')` — models/basic_modules.py:179-199.
    forward(pos, x, edge_index) -> [N, H + Dp]."""

    def __init__(self, nn_x: nn.Sequential, nn_pos: nn.Sequential, aggr: str = "max", **kwargs):
        super().__init__()
        if aggr != "max":
            raise NotImplementedError("only aggr='max' is used by the rigging networks")
        self.nn_x = nn_x
        self.nn_pos = nn_pos

    def forward(self, pos: torch.Tensor, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        self._guard(pos, x, edge_index)
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        pos = _lib.require_cuda(pos, "pos")
        x = _lib.require_cuda(x, "x")
        n, dev = x.shape[0], x.device
        if self.training:                            # train-mode BatchNorm + autograd (train_forward.py)
            from . import autograd_ops, train_forward
            g = train_forward.graph_for_training(self._graphs, edge_index, n)
            return autograd_ops.ConcatCols.apply(*train_forward.edge_conv_motion_parts(self, pos, x, g))

        def build():
            sd = self._device_state()
            wp_x, wq_x, b_x, br_x = packing._edge_mlp_parts(sd, "nn_x")
            wp_p, wq_p, b_p, br_p = packing._edge_mlp_parts(sd, "nn_pos")
            pq_x = packing.DenseLayer(W=packing._pack_wt(torch.cat([wp_x, wq_x])), K=wp_x.shape[1], N=2 * br_x.H,
                                      bias=packing._vec(torch.cat([b_x, torch.zeros_like(b_x)])))
            pq_p = packing.DenseLayer(W=packing._pack_wt(torch.cat([wp_p, wq_p])), K=wp_p.shape[1], N=2 * br_p.H,
                                      bias=packing._vec(torch.cat([b_p, torch.zeros_like(b_p)])))
            return pq_x, br_x, pq_p, br_p

        pq_x, br_x, pq_p, br_p = self._packed_for("edge", build)
        _check_rows("EdgeConvMotion", n, pos=pos)
        _check_width("EdgeConvMotion", x=(x, pq_x.K), pos=(pos, pq_p.K))
        g = self._graphs.get(edge_index, n)
        H, Dp = br_x.H, br_p.H
        out = torch.empty(n, H + Dp, device=dev, dtype=torch.float32)
        with engine.forward_scope(self._ws, dev):
            engine.fill(out, engine.NEG_INF)
            for layer, br, src, off in ((pq_x, br_x, x, 0), (pq_p, br_p, pos, H)):
                buf = self._ws.get(f"pq{off}", (n, layer.N), dev)
                engine.dense(layer, src, 0, src.shape[1], n, C=buf, ldc=layer.N)
                engine.edgeconv(br, buf, layer.N, 0, br.H, g, 1, out, H + Dp, off)
        return out


class GCUMotion(FusedModule):
    """`GCUMotion(in_channels, out_channels, in_channel_pos=3, dim_pos_feat=16, aggr='max')` —
    models/basic_modules.py:205-219.  forward(pos, x, tpl_edge_index, geo_edge_index) -> [N, out]."""

    def __init__(self, in_channels: int, out_channels: int, in_channel_pos: int = 3, dim_pos_feat: int = 16,
                 aggr: str = "max"):
        super().__init__()
        half = out_channels // 2
        self.edge_conv_tpl = EdgeConvMotion(nn_x=MLP([in_channels * 2, half, half]),
                                            nn_pos=MLP([in_channel_pos * 2, dim_pos_feat, dim_pos_feat]), aggr=aggr)
        self.edge_conv_geo = EdgeConvMotion(nn_x=MLP([in_channels * 2, half, half]),
                                            nn_pos=MLP([in_channel_pos * 2, dim_pos_feat, dim_pos_feat]), aggr=aggr)
        self.mlp = MLP([out_channels + dim_pos_feat * 2, out_channels])

    def forward(self, pos, x, tpl_edge_index, geo_edge_index) -> torch.Tensor:
        self._guard(pos, x, tpl_edge_index, geo_edge_index)
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        pos = _lib.require_cuda(pos, "pos")
        x = _lib.require_cuda(x, "x")
        n, dev = x.shape[0], x.device
        if self.training:
            from . import train_forward
            return train_forward.gcu_motion(self, pos, x, train_forward.graph_for_training(self._graphs, tpl_edge_index, n),
                                            train_forward.graph_for_training(self._graphs, geo_edge_index, n))

        def build():
            parts: list = []
            sd ={"g." + k: v for k, v in self._device_state().items()}
            gp = packing.pack_gcu(sd, "g", parts)
            return gp, packing.fuse_pos_pq(parts)

        gp, pq_pos = self._packed_for("gcu", build)
        _check_rows("GCUMotion", n, pos=pos)
        _check_width("GCUMotion", x=(x, gp.pq_x.K), pos=(pos, pq_pos.K))
        gt = self._graphs.get(tpl_edge_index, n)
        gg = self._graphs.get(geo_edge_index, n)
        pqpos = self._ws.get("pqpos", (n, pq_pos.N), dev)
        out = torch.empty(n, gp.out, device=dev, dtype=torch.float32)
        with engine.forward_scope(self._ws, dev):
            engine.dense(pq_pos, pos, 0, pos.shape[1], n, C=pqpos, ldc=pq_pos.N)
            engine.run_pos_branches(self._ws, ["gcu"], [gp], pqpos, gt, gg, n, 1)
            engine.run_gcu(self._ws, "gcu", gp, x, 0, x.shape[1], x.shape[1], gt, gg, n, 1, out, 0, gp.out)
        return out


class EdgeConv(FusedModule):
    """`EdgeConv(nn_pos, aggr='max')` — models/basic_modules.py:142-162: the EdgeConvMotion kernel without a second
    branch (`nn_pos` is the only MLP and is applied to x).  forward(x, edge_index) -> [N, H].
    Not reached from models/rignet.py (only models/corrnet.py:17-20); provided because it is the C_p = 0 case of
    the same fused kernel."""

    def __init__(self, nn_pos: nn.Sequential, aggr: str = "max", **kwargs):
        super().__init__()
        if aggr != "max":
            raise NotImplementedError("only aggr='max' is used by the reference networks")
        self.nn_pos = nn_pos

    def _pack(self):
        wp, wq, b, br = packing._edge_mlp_parts(self._device_state(), "nn_pos")
        w = torch.cat([wp, wq])
        pq = packing.DenseLayer(W=packing._pack_wt(w), K=wp.shape[1], N=2 * br.H,
                                bias=packing._vec(torch.cat([b, torch.zeros_like(b)]))).with_tc(w)
        return pq, br

    def run(self, x: torch.Tensor, g: engine.Graph, out: torch.Tensor, ldo: int, out_off: int) -> None:
        pq, br = self._packed_for("edge", self._pack)
        _check_width("EdgeConv", x=(x, pq.K))
        n = x.shape[0]
        buf = self._ws.get("pq", (n, pq.N), x.device)
        engine.dense(pq, x, 0, x.shape[1], n, C=buf, ldc=pq.N)
        engine.edgeconv(br, buf, pq.N, 0, br.H, g, 1, out, ldo, out_off)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        self._guard(x, edge_index)
        x = _lib.require_cuda(x.unsqueeze(-1) if x.dim() == 1 else x, "x")
        n = x.shape[0]
        if self.training:
            from . import train_forward
            return train_forward.edge_branch(x, self.nn_pos, train_forward.graph_for_training(self._graphs, edge_index, n))
        pq, br = self._packed_for("edge", self._pack)
        _check_width("EdgeConv", x=(x, pq.K))
        out = torch.empty(n, br.H, device=x.device, dtype=torch.float32)
        with engine.forward_scope(self._ws, x.device):
            engine.fill(out, engine.NEG_INF)
            self.run(x, self._graphs.get(edge_index, n), out, br.H, 0)
        return out


class GCU(FusedModule):
    """`GCU(in_channels, out_channels, aggr='max')` — models/basic_modules.py:165-177.
    forward(pos, tpl_edge_index, geo_edge_index) -> [N, out_channels] (its `pos` argument is the feature)."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "max"):
        super().__init__()
        half = out_channels // 2
        self.edge_conv_tpl = EdgeConv(nn_pos=MLP([in_channels * 2, half, half]), aggr=aggr)
        self.edge_conv_geo = EdgeConv(nn_pos=MLP([in_channels * 2, half, half]), aggr=aggr)
        self.mlp = MLP([out_channels, out_channels])

    def forward(self, pos, tpl_edge_index, geo_edge_index) -> torch.Tensor:
        self._guard(pos, tpl_edge_index, geo_edge_index)
        x = _lib.require_cuda(pos.unsqueeze(-1) if pos.dim() == 1 else pos, "pos")
        n, dev = x.shape[0], x.device
        if self.training:
            from . import train_forward
            return train_forward.gcu(self, x, train_forward.graph_for_training(self._graphs, tpl_edge_index, n),
                                     train_forward.graph_for_training(self._graphs, geo_edge_index, n))
        mlp = self._packed_for("gcu", lambda: packing.pack_mlp_layer(
            {"m." + k: v for k, v in self.mlp.state_dict(keep_vars=True).items()}, "m.0"))
        half = mlp.K // 2
        ec = self._ws.get("ec", (n, 2 * half), dev)
        out = torch.empty(n, mlp.N, device=dev, dtype=torch.float32)
        with engine.forward_scope(self._ws, dev):
            engine.fill(ec, engine.NEG_INF)
            self.edge_conv_tpl.run(x, self._graphs.get(tpl_edge_index, n), ec, 2 * half, 0)
            self.edge_conv_geo.run(x, self._graphs.get(geo_edge_index, n), ec, 2 * half, half)
            engine.dense(mlp, ec, 0, 2 * half, n, C=out, ldc=mlp.N)
        return out
