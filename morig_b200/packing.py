"""Folding and packing of reference-keyed parameters into the layouts the CUDA kernels read.

Everything here is algebra on parameters only (done once per weight version, in fp64, rounded
once to fp32), exact in real arithmetic:

* eval BatchNorm -> per-channel affine  y = s*r + t,  s = gamma / sqrt(var + 1e-5),  t = beta - mean*s
  (`MLP` = Linear -> ReLU -> BatchNorm, models/basic_modules.py:31-36: BN sits AFTER the ReLU);
* edge MLP layer 0 on cat[x_i, x_j - x_i] (models/basic_modules.py:193-194) with W0 = [Wa | Wb]:
      W0 [x_i, x_j - x_i] + b0 = (Wa - Wb) x_i + b0  +  Wb x_j  =  P[i] + Q[j]
  so P and Q are computed per VERTEX by one dense layer and the per-edge work starts at the ReLU;
* BN #1 of the edge MLP folds into its second Linear:  W1' = W1 diag(s0),  b1' = b1 + W1 t0;
  BN #2 stays an explicit affine applied before the max (s may be negative);
* GCNRig.mlp_transform layer 0 (models/rignet.py:65-66) reads cat[x_global, pos, feature, x1, x2, x3];
  its x_global block is constant per graph, so it becomes a per-graph bias (one tiny dense layer)
  and the remaining columns are re-ordered to this package's feature-buffer layout
  [x1 | x2 | x3 | feature | pos | 0-pad];
* TemporalAttn (models/rignet.py:36-46) only uses the cls row of the output, so q is the constant
  Wq cls and the per-head maps collapse to u_h = Wk_h^T q_h / sqrt(d), Mv_h = Wo[:, h] Wv_h, ...

All packed weights are stored transposed, [K, ldw] row-major with ldw a multiple of 4.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

BN_EPS = 1e-5


def _r4(n: int) -> int:
    return (n + 3) // 4 * 4


def _f64(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float64)


def bn_affine(sd, prefix: str):
    """(s, t) in fp64 of the eval BatchNorm stored under `prefix` (.weight/.bias/.running_*)."""
    s = _f64(sd[prefix + ".weight"]) / torch.sqrt(_f64(sd[prefix + ".running_var"]) + BN_EPS)
    t = _f64(sd[prefix + ".bias"]) - _f64(sd[prefix + ".running_mean"]) * s
    return s, t


def _pack_wt(w_out_in: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """[out, in] fp64 -> contiguous fp32 [k_pad, ldw] = transposed, zero padded."""
    out_c, in_c = w_out_in.shape
    k = in_c if k_pad is None else k_pad
    buf = torch.zeros(k, _r4(out_c), dtype=torch.float64, device=w_out_in.device)
    buf[:in_c, :out_c] = w_out_in.t()
    return buf.to(torch.float32).contiguous()


def _vec(v: torch.Tensor) -> torch.Tensor:
    return v.to(torch.float32).contiguous()


def tf32_rna(x32: torch.Tensor) -> torch.Tensor:
    """round-to-nearest (ties away) of fp32 to TF32 (10 explicit mantissa bits), like cvt.rna.tf32.f32"""
    bits = x32.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def tc_tile_n(n: int) -> int:
    """n-tile (output channels per tile = 128 TMEM lanes per accumulator) used by the tcgen05 engine for a layer
    with n outputs; narrower layers are zero-padded to 128 channels"""
    return 128 if n <= 128 else 256


KIND_TF32, KIND_F16 = 0, 1


def tc_kind() -> int:
    """operand kind of the tcgen05 engine: fp16 split (default) or tf32 split (`MORIG_TC_KIND=tf32`)"""
    v = os.environ.get("MORIG_TC_KIND", "f16").lower()
    if v not in ("f16", "tf32"):
        raise ValueError(f"MORIG_TC_KIND={v!r}: expected 'f16' or 'tf32'")
    return KIND_F16 if v == "f16" else KIND_TF32


def pack_tc_blob(w_out_in: torch.Tensor, k: int, bn: int, kind: int = KIND_TF32):
    """Tensor-core image of a Linear weight [N, K'] (fp64): split W ~ hi + lo (the kernel issues
    A_lo*B_hi + A_hi*B_lo + A_hi*B_hi) and laid out as the byte image of the kernel's shared-memory B stage:

        blob[n_tile][k_chunk][half = hi|lo][row n (bn)][16-byte chunk c ^ (n % 8)][4 tf32 | 8 fp16]

    i.e. K-major rows of 128 bytes (32 tf32 / 64 fp16) with the SWIZZLE_128B pattern already applied, so that one
    cp.async.bulk per (n_tile, k_chunk) fills the stage (csrc/gemm_tc.cuh).  K is zero-padded to a multiple of
    the chunk (`k` = the K the kernel will be launched with), N to a multiple of bn.

    kind TF32: hi = rna_tf32(W), lo = rna_tf32(W - hi).
    kind F16:  fp16 has a 5-bit exponent, so the image holds W * 2^j with j chosen such that max|W| * 2^j lies in
               [2^14, 2^15); hi = fp16(W 2^j), lo = fp16(W 2^j - hi).  Elements down to 2^-17 of max|W| keep 22
               significant bits.  Returns (blob, 2^-j); the kernel's epilogue multiplies by it.
    Returns (blob, w_inv) with w_inv = 1.0 for TF32."""
    n, kin = w_out_in.shape
    kc = 64 if kind == KIND_F16 else 32
    per = kc // 8                                                     # elements per 16-byte chunk
    nk = (k + kc - 1) // kc
    nt = (n + bn - 1) // bn
    wp = torch.zeros(nt * bn, nk * kc, dtype=torch.float64, device=w_out_in.device)
    wp[:n, :kin] = w_out_in
    w_inv = 1.0
    if kind == KIND_F16:
        amax = float(wp.abs().max())
        j = 14 - math.floor(math.log2(amax)) if amax > 0.0 and math.isfinite(amax) else 0
        j = max(-100, min(100, j))
        ws = wp * (2.0 ** j)
        hi = ws.to(torch.float16)
        lo = (ws - hi.to(torch.float64)).to(torch.float16)
        w_inv = 2.0 ** (-j)
    else:
        hi = tf32_rna(wp.to(torch.float32))
        lo = tf32_rna((wp - hi.to(torch.float64)).to(torch.float32))
    both = torch.stack([hi, lo], dim=0)                               # [2, nt*bn, nk*kc]
    both = both.reshape(2, nt, bn, nk, 8, per).permute(1, 3, 0, 2, 4, 5)  # [nt, nk, 2, bn, chunk, per]
    rows = torch.arange(bn, device=wp.device)
    src_chunk = torch.arange(8, device=wp.device).unsqueeze(0) ^ (rows % 8).unsqueeze(1)   # dst chunk d holds src d^(n%8)
    idx = src_chunk.view(1, 1, 1, bn, 8, 1).expand(nt, nk, 2, bn, 8, per)
    return torch.gather(both, 4, idx).contiguous().reshape(-1), w_inv


@dataclass
class DenseLayer:
    """one `morig_dense_fwd` worth of parameters"""
    W: torch.Tensor            # [K, ldw]
    K: int
    N: int
    bias: Optional[torch.Tensor] = None
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    relu: bool = False
    Wtc: Optional[torch.Tensor] = None     # tcgen05 image (pack_tc_blob) or None -> CUDA-core engine
    tc_bn: int = 0
    tc_kind: int = KIND_TF32
    tc_w_inv: float = 1.0

    @property
    def ldw(self) -> int:
        return self.W.shape[1]

    def with_tc(self, w_out_in: torch.Tensor, kind: Optional[int] = None) -> "DenseLayer":
        """attach the tensor-core image when the layer is large enough to benefit (K >= 32, N >= 16; narrow outputs are
        zero-padded to one 128-channel tile: these layers are bound by their rows, not by the MMAs)"""
        if self.K >= 32 and self.K % 4 == 0 and self.N >= 16:
            self.tc_bn = tc_tile_n(self.N)
            self.tc_kind = tc_kind() if kind is None else kind
            self.Wtc, self.tc_w_inv = pack_tc_blob(w_out_in, self.K, self.tc_bn, self.tc_kind)
        return self


def pack_mlp_layer(sd, prefix: str, k_pad: Optional[int] = None, cols: Optional[torch.Tensor] = None) -> DenseLayer:
    """`{prefix}.0` Linear + ReLU + `{prefix}.2` BatchNorm as one fused dense layer.
    `cols` optionally selects / re-orders the Linear's input columns."""
    w = _f64(sd[prefix + ".0.weight"])
    if cols is not None:
        w = w[:, cols]
    s, t = bn_affine(sd, prefix + ".2")
    return DenseLayer(W=_pack_wt(w, k_pad), K=(w.shape[1] if k_pad is None else k_pad), N=w.shape[0],
                      bias=_vec(_f64(sd[prefix + ".0.bias"])), scale=_vec(s), shift=_vec(t), relu=True).with_tc(w)


def pack_linear(sd, prefix: str, bias: bool = True) -> DenseLayer:
    w = _f64(sd[prefix + ".weight"])
    return DenseLayer(W=_pack_wt(w), K=w.shape[1], N=w.shape[0],
                      bias=_vec(_f64(sd[prefix + ".bias"])) if bias else None).with_tc(w)


@dataclass
class EdgeBranch:
    """second half of one edge MLP (per edge set): h -> relu(h W1' + b1') * s1 + t1"""
    W1: torch.Tensor
    b1: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor
    H: int
    W1tc: Optional[torch.Tensor] = None    # tcgen05 image of W1 (H >= 64)
    tc_kind: int = KIND_TF32
    tc_w_inv: float = 1.0


def _edge_mlp_parts(sd, prefix: str):
    """split a 2-layer edge MLP `prefix` (= ...nn_x or ...nn_pos) into the per-vertex PQ block and the
    per-edge branch. Returns (Wp [H,C], Wq [H,C], bp [H], EdgeBranch)."""
    w0 = _f64(sd[prefix + ".0.0.weight"])             # [H, 2C]
    b0 = _f64(sd[prefix + ".0.0.bias"])
    c = w0.shape[1] // 2
    wa, wb = w0[:, :c], w0[:, c:]
    s0, t0 = bn_affine(sd, prefix + ".0.2")
    w1 = _f64(sd[prefix + ".1.0.weight"])             # [H, H]
    b1 = _f64(sd[prefix + ".1.0.bias"])
    s1, t1 = bn_affine(sd, prefix + ".1.2")
    w1f = w1 * s0.unsqueeze(0)
    b1f = b1 + w1 @ t0
    br = EdgeBranch(W1=_pack_wt(w1f), b1=_vec(b1f), scale=_vec(s1), shift=_vec(t1), H=w1.shape[0])
    if br.H in (64, 128, 256):
        br.tc_kind = tc_kind()
        br.W1tc, br.tc_w_inv = pack_edge_tc_blob(w1f, br.scale, br.tc_kind)
    return wa - wb, wb, b0, br


def pack_edge_tc_blob(w1f: torch.Tensor, scale32: torch.Tensor, kind: int):
    """Tensor-core image of the second edge Linear for the fused EdgeConv kernel.  The kernel takes the segment's
    maximum on the RAW accumulators and applies bias -> ReLU -> BatchNorm affine once per segment; that is exact because
    the chain is monotone in sigma * z with sigma = sign of the BatchNorm scale (for a negative scale the largest output
    comes from the smallest pre-activation), so the rows of channels with a negative (fp32) scale are negated here and
    the kernel multiplies the extreme by sigma again (csrc/gemm_tc.cuh, EPI_SEGMAX flush)."""
    sigma = torch.where(scale32.to(w1f.device) < 0, -1.0, 1.0).to(torch.float64)
    return pack_tc_blob(w1f * sigma.unsqueeze(1), w1f.shape[0], tc_tile_n(w1f.shape[0]), kind)


@dataclass
class GCUPack:
    """one GCUMotion: x-branch PQ layer (both edge sets), four edge branches, the vertex mlp"""
    pq_x: DenseLayer                       # [C_x(pad4), 4H]: columns [P_tpl | Q_tpl | P_geo | Q_geo]
    x_tpl: EdgeBranch
    x_geo: EdgeBranch
    pos_tpl: EdgeBranch
    pos_geo: EdgeBranch
    pos_col: int                           # column of this GCU's block inside the shared pos PQ buffer
    mlp: DenseLayer
    H: int
    Dp: int
    out: int


def pack_gcu(sd, prefix: str, pos_parts: list, k_pad_x: Optional[int] = None) -> GCUPack:
    """`pos_parts` collects (Wp, Wq, bp) of the pos branches so the caller can fuse all GCUs' pos PQ
    layers (they share the same input) into one dense layer."""
    px_t = _edge_mlp_parts(sd, prefix + ".edge_conv_tpl.nn_x")
    px_g = _edge_mlp_parts(sd, prefix + ".edge_conv_geo.nn_x")
    pp_t = _edge_mlp_parts(sd, prefix + ".edge_conv_tpl.nn_pos")
    pp_g = _edge_mlp_parts(sd, prefix + ".edge_conv_geo.nn_pos")
    H, Dp = px_t[3].H, pp_t[3].H
    w = torch.cat([px_t[0], px_t[1], px_g[0], px_g[1]], dim=0)              # [4H, C]
    b = torch.cat([px_t[2], torch.zeros_like(px_t[2]), px_g[2], torch.zeros_like(px_g[2])])
    pq_x = DenseLayer(W=_pack_wt(w, k_pad_x), K=(w.shape[1] if k_pad_x is None else k_pad_x), N=4 * H,
                      bias=_vec(b)).with_tc(w)
    pos_col = sum(p[0].shape[0] for p in pos_parts)
    for part in (pp_t, pp_g):
        pos_parts.append((part[0], part[2]))                               # P block (+bias)
        pos_parts.append((part[1], torch.zeros_like(part[2])))             # Q block
    mlp = pack_mlp_layer(sd, prefix + ".mlp.0")
    return GCUPack(pq_x=pq_x, x_tpl=px_t[3], x_geo=px_g[3], pos_tpl=pp_t[3], pos_geo=pp_g[3], pos_col=pos_col,
                   mlp=mlp, H=H, Dp=Dp, out=mlp.N)


def fuse_pos_pq(pos_parts: list, k_pad: Optional[int] = None) -> DenseLayer:
    w = torch.cat([p[0] for p in pos_parts], dim=0)
    b = torch.cat([p[1] for p in pos_parts])
    return DenseLayer(W=_pack_wt(w, k_pad), K=(w.shape[1] if k_pad is None else k_pad), N=w.shape[0],
                      bias=_vec(b)).with_tc(w)


@dataclass
class GCNRigPack:
    F: int                                 # feature channels
    O: int                                 # output channels
    gcus: List[GCUPack]
    pq_pos: DenseLayer                     # fused pos PQ of the three GCUs: [3, 3*4*Dp]
    glb: DenseLayer                        # mlp_glb 832 -> 1024 (pooled, never stored)
    t0_global: DenseLayer                  # x_global block of mlp_transform.0.0 -> per-graph bias
    t0: DenseLayer                         # remaining columns in feature-buffer order
    t1: DenseLayer
    head: DenseLayer
    ldf: int                               # row stride of the feature buffer
    x_off: List[int] = field(default_factory=lambda: [0, 64, 320])
    feat_off: int = 832
    pos_off: int = 0


def pack_gcn_rig(sd: Dict[str, torch.Tensor], prefix: str) -> GCNRigPack:
    """GCNRig (models/rignet.py:50-67) under state_dict prefix `prefix` (e.g. 'motionNet')."""
    F = sd[prefix + ".gcu_1.edge_conv_tpl.nn_x.0.0.weight"].shape[1] // 2
    pos_parts: list = []
    g1 = pack_gcu(sd, prefix + ".gcu_1", pos_parts, k_pad_x=_r4(F))
    g2 = pack_gcu(sd, prefix + ".gcu_2", pos_parts)
    g3 = pack_gcu(sd, prefix + ".gcu_3", pos_parts)
    xw = g1.out + g2.out + g3.out                                           # 832
    glb = pack_mlp_layer(sd, prefix + ".mlp_glb.0")
    G = glb.N                                                               # 1024
    # mlp_transform.0.0 input columns: [x_global (G) | pos (3) | feature (F) | x1 | x2 | x3]
    w_t0 = _f64(sd[prefix + ".mlp_transform.0.0.0.weight"])
    t0_global = DenseLayer(W=_pack_wt(w_t0[:, :G]), K=G, N=w_t0.shape[0]).with_tc(w_t0[:, :G])
    dev = w_t0.device
    cols = torch.cat([torch.arange(G + 3 + F, G + 3 + F + xw, device=dev),  # x1|x2|x3
                      torch.arange(G + 3, G + 3 + F, device=dev),           # feature
                      torch.arange(G, G + 3, device=dev)])                  # pos
    k_t0 = _r4(xw + F + 3)
    t0 = pack_mlp_layer(sd, prefix + ".mlp_transform.0.0", k_pad=k_t0, cols=cols)
    t1 = pack_mlp_layer(sd, prefix + ".mlp_transform.0.1")
    head = pack_linear(sd, prefix + ".mlp_transform.1")
    ldf = xw + (F + 3 + 31) // 32 * 32
    return GCNRigPack(F=F, O=head.N, gcus=[g1, g2, g3], pq_pos=fuse_pos_pq(pos_parts), glb=glb,
                      t0_global=t0_global, t0=t0, t1=t1, head=head, ldf=ldf,
                      x_off=[0, g1.out, g1.out + g2.out], feat_off=xw, pos_off=xw + F)


@dataclass
class AttnPack:
    C: int
    heads: int
    D: int
    u: torch.Tensor          # [heads, C]
    l0: torch.Tensor         # [heads]
    Mv: torch.Tensor         # [heads, D, C]
    c0: torch.Tensor         # [heads, D]
    ff0: DenseLayer
    ff1: DenseLayer


def pack_temporal_attn(sd, prefix: str, num_heads: int = 2) -> AttnPack:
    """TemporalAttn (models/rignet.py:10-46) collapsed onto the cls query row."""
    wq, wk, wv = (_f64(sd[f"{prefix}.w_{n}s.weight"]) for n in "qkv")       # [heads*d, C]
    wo = _f64(sd[prefix + ".w_o.weight"])                                   # [D, heads*d]
    cls = _f64(sd[prefix + ".cls_token"]).reshape(-1)                       # [C]
    d = wq.shape[0] // num_heads
    q, kc, vc = wq @ cls, wk @ cls, wv @ cls
    u, l0, mv, c0 = [], [], [], []
    for h in range(num_heads):
        sl = slice(h * d, (h + 1) * d)
        u.append(wk[sl].t() @ q[sl] / math.sqrt(d))
        l0.append((q[sl] @ kc[sl]) / math.sqrt(d))
        mv.append(wo[:, sl] @ wv[sl])
        c0.append(wo[:, sl] @ vc[sl])
    return AttnPack(C=wq.shape[1], heads=num_heads, D=wo.shape[0],
                    u=_vec(torch.stack(u)), l0=_vec(torch.stack(l0)), Mv=_vec(torch.stack(mv)),
                    c0=_vec(torch.stack(c0)),
                    ff0=pack_mlp_layer(sd, prefix + ".feedforward.0"),
                    ff1=pack_mlp_layer(sd, prefix + ".feedforward.1"))


@dataclass
class SkinPack:
    """SkinNet_inner (models/rignet.py:136-182)"""
    in_pos: int                            # 3 + selected skin columns
    k_pos: int                             # in_pos rounded up to 4 (row stride of the raw-input buffer)
    skin_cols: torch.Tensor                # int32 selected columns of data.skin_input
    gcus: List[GCUPack]
    pq_pos: DenseLayer
    g0: DenseLayer                         # multi_layer_tranform2.0: 256 -> 512
    g1: DenseLayer                         # multi_layer_tranform2.1: 512 -> 1024 (pooled)
    c0_global: DenseLayer                  # x_global block of cls_branch.0.0 -> per-graph bias
    c0: DenseLayer                         # x_3 block
    c1: DenseLayer
    head: DenseLayer


def skin_columns(width: int, nearest_bone: int, use_Dg: bool, use_Lf: bool) -> List[int]:
    """Columns of `data.skin_input` kept by the selection at models/rignet.py:159-171."""
    cols = list(range(width))
    if use_Dg and use_Lf:
        return cols[: 8 * nearest_bone]
    if use_Dg and not use_Lf:
        return [c for c in cols if c % 8 != 7][: 7 * nearest_bone]
    if use_Lf and not use_Dg:
        return [c for c in cols if c % 8 != 6][: 7 * nearest_bone]
    cols = [c for c in cols if c % 8 != 7]
    cols = [c for i, c in enumerate(cols) if i % 7 != 6]
    return cols[: 6 * nearest_bone]


def pack_skin(sd, prefix: str, skin_width: int, nearest_bone: int, use_Dg: bool, use_Lf: bool) -> SkinPack:
    dev = sd[prefix + ".gcu1.mlp.0.0.weight"].device
    cols = skin_columns(skin_width, nearest_bone, use_Dg, use_Lf)
    in_pos = 3 + len(cols)
    if sd[prefix + ".gcu1.edge_conv_tpl.nn_pos.0.0.weight"].shape[1] != 2 * in_pos:
        raise ValueError("skin_input width does not match the network's in_channel_pos")
    k_pos = _r4(in_pos)
    pos_parts: list = []
    gcus = [pack_gcu(sd, f"{prefix}.gcu{i}", pos_parts) for i in (1, 2, 3)]
    w_c0 = _f64(sd[prefix + ".cls_branch.0.0.0.weight"])                    # [1024, 256 + 1024]
    c_x = gcus[2].out
    c0 = pack_mlp_layer(sd, prefix + ".cls_branch.0.0", cols=torch.arange(0, c_x, device=dev))
    c0_global = DenseLayer(W=_pack_wt(w_c0[:, c_x:]), K=w_c0.shape[1] - c_x, N=w_c0.shape[0]).with_tc(w_c0[:, c_x:])
    return SkinPack(in_pos=in_pos, k_pos=k_pos, skin_cols=torch.tensor(cols, dtype=torch.int32, device=dev),
                    gcus=gcus, pq_pos=fuse_pos_pq(pos_parts, k_pad=k_pos),
                    g0=pack_mlp_layer(sd, prefix + ".multi_layer_tranform2.0"),
                    g1=pack_mlp_layer(sd, prefix + ".multi_layer_tranform2.1"),
                    c0_global=c0_global, c0=c0,
                    c1=pack_mlp_layer(sd, prefix + ".cls_branch.0.1"),
                    head=pack_linear(sd, prefix + ".cls_branch.1"))
