"""Launch helpers of the training path: torch tensors in, torch tensors out, all arithmetic in
`libmorig_b200.so` (csrc/train.cu + the fp32 dense engine).  `autograd_ops.py` builds the autograd functions on top.

Every matrix argument is a 2-D fp32 CUDA tensor whose last dimension is contiguous; row strides are passed to the
kernels, so column slices of wider buffers (the P / Q halves of a factorised edge layer, key-frame slices of the flow)
are used in place.  Outputs are freshly allocated by torch (the caching allocator makes that a pointer bump).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

BN_EPS = 1e-5


def _r4(n: int) -> int:
    return (n + 3) // 4 * 4


def mat(t: torch.Tensor) -> torch.Tensor:
    """2-D fp32 view usable by the kernels (last dim contiguous, non-negative row stride); copies otherwise"""
    if t.dim() != 2:
        raise ValueError(f"expected a matrix, got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    ok = (t.shape[1] == 1 or t.stride(1) == 1) and (t.shape[0] == 1 or t.stride(0) >= t.shape[1])
    return t if ok else t.contiguous()


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _sp() -> int:
    return _lib.stream_ptr()


# ---- operand ranges for the fp16-split tensor-core GEMMs -------------------------------------------------------------------
# The kernels that produce the big activation / gradient matrices (BatchNorm apply, BatchNorm-ReLU backward, edge gather)
# also reduce max |output| into a device scalar, which travels with the tensor object as `_morig_amax` =
# (scalar, data_ptr, version).  A consumer only trusts it for exactly that tensor in exactly that state; anything else
# (views, copies, in-place edits, tensors from elsewhere) falls back to the |max| pass of `engine._amax_in`.
_amax_pool: dict = {}                    # device -> [zeroed block, cursor]: one memset per 4096 scalars, not one per tensor


def _new_amax(dev) -> torch.Tensor:
    """a zeroed device scalar; slices of a block that stays alive as long as any tag refers to it and is never reused"""
    ent = _amax_pool.get(dev)
    if ent is None or ent[1] >= ent[0].numel():
        ent = _amax_pool[dev] = [torch.zeros(4096, dtype=torch.float32, device=dev), 0]
    i = ent[1]
    ent[1] = i + 1
    return ent[0][i:i + 1]


def _tag_amax(t: torch.Tensor, amax: torch.Tensor) -> None:
    t._morig_amax = (amax, t.data_ptr(), t._version)


def _known_amax(t: torch.Tensor) -> int:
    tag = getattr(t, "_morig_amax", None)
    if tag is not None and tag[1] == t.data_ptr() and tag[2] == t._version and tag[0].device == t.device:
        return tag[0].data_ptr()
    return 0


def _tc_ok(M: int, K: int, N: int, A: torch.Tensor) -> bool:
    """shapes the tcgen05 split-fp16 engine takes (csrc/gemm_tc.cuh); MORIG_TRAIN_TC=0 keeps training on the fp32 engine"""
    import os
    return (os.environ.get("MORIG_TRAIN_TC", "1") != "0" and M >= 256 and K >= 32 and K % 4 == 0 and N >= 16
            and _ld(A) % 4 == 0 and A.data_ptr() % 16 == 0)


def dense_tc(A: torch.Tensor, w: torch.Tensor, transposed: bool, n_out: int, k: int, bias: Optional[torch.Tensor],
             relu: bool) -> torch.Tensor:
    """C [M, n_out] = act(A [M, k] @ Wl^T + bias) on the tensor-core engine, Wl [n_out, k] = w (or w^T when `transposed`);
    the fp16-split weight image is packed on the device (the weights change every optimisation step)"""
    from . import engine, packing
    lib = _lib.load()
    known = _known_amax(A)
    A0 = A
    A = mat(A)
    if A is not A0:
        known = 0
    M = A.shape[0]
    dev = A.device
    bn = packing.tc_tile_n(n_out)
    w = mat(w)
    blob = torch.empty(lib.morig_pack_tc_f16_bytes(n_out, k, bn), dtype=torch.uint8, device=dev)
    scal = torch.empty(2, dtype=torch.float32, device=dev)                     # [w_inv, amax scratch]
    _lib.check(lib.morig_pack_tc_f16(w.data_ptr(), _ld(w), n_out, k, 1 if transposed else 0, bn, blob.data_ptr(), scal.data_ptr(),
                                     scal.data_ptr() + 4, _sp()), "morig_pack_tc_f16")
    C_ = torch.empty(M, n_out, dtype=torch.float32, device=dev)
    d = _lib.DenseDesc()
    d.A, d.lda = A.data_ptr(), _ld(A)
    d.W, d.ldw = blob.data_ptr(), _r4(n_out)                                    # (unused by the tensor-core path; must be valid)
    d.bias = _lib.ptr(bias)
    d.C, d.ldc = C_.data_ptr(), n_out
    d.M, d.N, d.K = M, n_out, k
    d.relu = 1 if relu else 0
    d.Wtc, d.tc_bn, d.tc_kind, d.tc_w_inv = blob.data_ptr(), bn, packing.KIND_F16, 0.0
    d.tc_w_inv_dev = scal.data_ptr()
    d.a_amax = known if (known and A.shape[1] == k) else engine._amax_in(A, 0, _ld(A), M, k)
    _lib.check(lib.morig_dense_fwd(ctypes.byref(d), _sp()), "morig_dense_fwd")
    return C_


def dense_ffma(A: torch.Tensor, Wt: torch.Tensor, n_out: int, bias: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
    """C [M, n_out] = act(A [M, K] @ Wt [K, ldw] + bias) on the fp32 CUDA-core engine (morig_dense_fwd without a
    tensor-core image)"""
    A = mat(A)
    M, K = A.shape
    C_ = torch.empty(M, n_out, dtype=torch.float32, device=A.device)
    d = _lib.DenseDesc()
    d.A, d.lda = A.data_ptr(), _ld(A)
    d.W, d.ldw = Wt.data_ptr(), Wt.stride(0)
    d.bias = _lib.ptr(bias)
    d.C, d.ldc = C_.data_ptr(), n_out
    d.M, d.N, d.K = M, n_out, K
    d.relu = 1 if relu else 0
    _lib.check(_lib.load().morig_dense_fwd(ctypes.byref(d), _sp()), "morig_dense_fwd")
    return C_


def transpose_pad(w: torch.Tensor) -> torch.Tensor:
    """Linear weight [N, K] -> packed [K, r4(N)] operand of the dense engine"""
    w = mat(w)
    n, k = w.shape
    out = torch.empty(k, _r4(n), dtype=torch.float32, device=w.device)
    _lib.check(_lib.load().morig_transpose_pad_f32(w.data_ptr(), n, k, _ld(w), out.data_ptr(), out.stride(0), _sp()),
               "morig_transpose_pad_f32")
    return out


def linear_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = False) -> torch.Tensor:
    """act(x @ w^T + b), w [N, K] as torch stores a Linear weight"""
    x = mat(x)
    b = None if b is None else b.contiguous()
    if _tc_ok(x.shape[0], x.shape[1], w.shape[0], x):
        return dense_tc(x, w, False, w.shape[0], w.shape[1], b, relu)
    return dense_ffma(x, transpose_pad(w), w.shape[0], b, relu)


def matmul_nn(dy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """dy [M, N] @ w [N, K]: the input gradient of a Linear (the weight is used in its native layout)"""
    w = mat(w)
    n, k = w.shape
    dy = mat(dy)
    if _tc_ok(dy.shape[0], n, k, dy):
        return dense_tc(dy, w, True, k, n, None, False)
    if w.stride(0) % 4 != 0 or w.data_ptr() % 16 != 0:
        wp = torch.zeros(n, _r4(k), dtype=torch.float32, device=w.device)
        _lib.check(_lib.load().morig_gather_cols(w.data_ptr(), _ld(w), 0, 0, 0, k, n, 1, wp.data_ptr(), wp.stride(0), 0, 0,
                                                 _sp()), "morig_gather_cols")
        w = wp
    return dense_ffma(dy, w, k, None, False)


def wgrad(dy: torch.Tensor, x: torch.Tensor, want_bias: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """dW [N, K] = dy^T x, db [N] = column sums of dy"""
    lib = _lib.load()
    dy, x = mat(dy), mat(x)
    M, N = dy.shape
    K = x.shape[1]
    dW = torch.empty(N, K, dtype=torch.float32, device=dy.device)
    db = torch.empty(N, dtype=torch.float32, device=dy.device) if want_bias else None
    nbytes = lib.morig_wgrad_workspace(M, N, K)
    ws = _ws(nbytes, dy.device)
    _lib.check(lib.morig_wgrad_f32(dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), M, N, K, 0, 0, dW.data_ptr(), K,
                                   _lib.ptr(db), 0, ws.data_ptr(), nbytes, _sp()), "morig_wgrad_f32")
    return dW, db


def bn_train_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: Optional[torch.Tensor],
                 running_var: Optional[torch.Tensor], momentum: float, eps: float = BN_EPS):
    """train-mode BatchNorm1d: returns (y, mean, invstd); running statistics are updated in place"""
    lib = _lib.load()
    x = mat(x)
    R, C = x.shape
    dev = x.device
    y = torch.empty(R, C, dtype=torch.float32, device=dev)
    stats = torch.empty(4, C, dtype=torch.float32, device=dev)          # mean, invstd, scale, shift
    amax = _new_amax(dev)
    nbytes = lib.morig_colstats_workspace(R, C)
    ws = _ws(nbytes, dev)
    _lib.check(lib.morig_bn_train_fwd(x.data_ptr(), _ld(x), R, C, gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                      _lib.ptr(running_mean), _lib.ptr(running_var), stats[0].data_ptr(),
                                      stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), y.data_ptr(), C,
                                      amax.data_ptr(), ws.data_ptr(), nbytes, _sp()), "morig_bn_train_fwd")
    _tag_amax(y, amax)
    return y, stats[0], stats[1]


def bn_relu_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor,
                relu: bool):
    """gradient at the input of [ReLU ->] BatchNorm(train): returns (dz, dgamma, dbeta)"""
    lib = _lib.load()
    dy, x = mat(dy), mat(x)
    R, C = x.shape
    dev = x.device
    dz = torch.empty(R, C, dtype=torch.float32, device=dev)
    dg = torch.empty(C, dtype=torch.float32, device=dev)
    db = torch.empty(C, dtype=torch.float32, device=dev)
    coef = torch.empty(3 * C, dtype=torch.float32, device=dev)
    amax = _new_amax(dev)
    nbytes = lib.morig_colstats_workspace(R, C)
    ws = _ws(nbytes, dev)
    _lib.check(lib.morig_bn_relu_bwd(dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), R, C, gamma.data_ptr(), mean.data_ptr(),
                                     invstd.data_ptr(), 1 if relu else 0, dz.data_ptr(), C, dg.data_ptr(), db.data_ptr(),
                                     coef.data_ptr(), amax.data_ptr(), ws.data_ptr(), nbytes, _sp()), "morig_bn_relu_bwd")
    _tag_amax(dz, amax)
    return dz, dg, db


def edge_gather_relu(P: torch.Tensor, Q: torch.Tensor, g) -> torch.Tensor:
    """h [E', C] = relu(P[tgt] + Q[col]) over the valid CSR slots of graph `g` (engine.Graph with e_real set)"""
    P, Q = mat(P), mat(Q)
    C = P.shape[1]
    h = torch.empty(g.e_real, C, dtype=torch.float32, device=P.device)
    amax = _new_amax(P.device)
    _lib.check(_lib.load().morig_edge_gather_relu(P.data_ptr(), _ld(P), Q.data_ptr(), _ld(Q), g.tgt.data_ptr(),
                                                  g.col.data_ptr(), g.e_real, C, h.data_ptr(), C, amax.data_ptr(), _sp()),
               "morig_edge_gather_relu")
    _tag_amax(h, amax)
    return h


def edge_gather_relu_bwd(dh: torch.Tensor, h: torch.Tensor, g) -> torch.Tensor:
    """returns dPQ [N, 2C] = [dP | dQ]"""
    dh, h = mat(dh), mat(h)
    C = h.shape[1]
    dpq = torch.empty(g.n, 2 * C, dtype=torch.float32, device=h.device)
    _lib.check(_lib.load().morig_edge_gather_relu_bwd(dh.data_ptr(), _ld(dh), h.data_ptr(), _ld(h), g.rowptr.data_ptr(),
                                                      g.col.data_ptr(), g.n, g.e_real, C, dpq.data_ptr(), 2 * C,
                                                      dpq.data_ptr() + 4 * C, 2 * C, _sp()), "morig_edge_gather_relu_bwd")
    return dpq


def segmax_fwd(y: torch.Tensor, ptr: torch.Tensor, S: int):
    y = mat(y)
    C = y.shape[1]
    out = torch.empty(S, C, dtype=torch.float32, device=y.device)
    arg = torch.empty(S, C, dtype=torch.int32, device=y.device)
    _lib.check(_lib.load().morig_segmax_fwd(y.data_ptr(), _ld(y), ptr.data_ptr(), S, C, out.data_ptr(), C, arg.data_ptr(), C,
                                            _sp()), "morig_segmax_fwd")
    return out, arg


def segmax_bwd(dout: torch.Tensor, arg: torch.Tensor, R: int) -> torch.Tensor:
    dout = mat(dout)
    S, C = arg.shape
    dy = torch.empty(R, C, dtype=torch.float32, device=dout.device)
    _lib.check(_lib.load().morig_segmax_bwd(dout.data_ptr(), _ld(dout), arg.data_ptr(), C, S, C, dy.data_ptr(), C, R, _sp()),
               "morig_segmax_bwd")
    return dy


def seg_ptr(keys32: torch.Tensor, S: int) -> torch.Tensor:
    ptr = torch.empty(S + 1, dtype=torch.int32, device=keys32.device)
    _lib.check(_lib.load().morig_seg_ptr(keys32.data_ptr(), keys32.numel(), S, ptr.data_ptr(), _sp()), "morig_seg_ptr")
    return ptr


def row_gather(src: torch.Tensor, idx32: torch.Tensor) -> torch.Tensor:
    src = mat(src)
    R, C = idx32.numel(), src.shape[1]
    out = torch.empty(R, C, dtype=torch.float32, device=src.device)
    _lib.check(_lib.load().morig_row_gather(src.data_ptr(), _ld(src), idx32.data_ptr(), R, C, out.data_ptr(), C, _sp()),
               "morig_row_gather")
    return out


def seg_sum(x: torch.Tensor, ptr: torch.Tensor, S: int) -> torch.Tensor:
    x = mat(x)
    C = x.shape[1]
    out = torch.empty(S, C, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().morig_seg_sum(x.data_ptr(), _ld(x), ptr.data_ptr(), S, C, out.data_ptr(), C, _sp()), "morig_seg_sum")
    return out


def normalize_fwd(x: torch.Tensor) -> torch.Tensor:
    x = mat(x)
    R, C = x.shape
    y = torch.empty(R, C, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().morig_normalize_fwd(x.data_ptr(), _ld(x), R, C, y.data_ptr(), C, _sp()), "morig_normalize_fwd")
    return y


def normalize_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    x, dy = mat(x), mat(dy)
    R, C = x.shape
    dx = torch.empty(R, C, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().morig_normalize_bwd(x.data_ptr(), _ld(x), dy.data_ptr(), _ld(dy), R, C, dx.data_ptr(), C, _sp()),
               "morig_normalize_bwd")
    return dx


def attn_cls_fwd(q0, kc, vc, Kx, Vx, d: int):
    """q0 / kc / vc [HD], Kx / Vx [N, T, HD] contiguous -> (out [N, HD], att [N, heads, T + 1])"""
    N, T, HD = Kx.shape
    out = torch.empty(N, HD, dtype=torch.float32, device=Kx.device)
    att = torch.empty(N, HD // d, T + 1, dtype=torch.float32, device=Kx.device)
    _lib.check(_lib.load().morig_attn_cls_fwd(q0.data_ptr(), kc.data_ptr(), vc.data_ptr(), Kx.data_ptr(), Vx.data_ptr(), N, T,
                                              HD, d, out.data_ptr(), att.data_ptr(), _sp()), "morig_attn_cls_fwd")
    return out, att


def attn_cls_bwd(q0, kc, vc, Kx, Vx, att, dout, d: int):
    lib = _lib.load()
    N, T, HD = Kx.shape
    dev = Kx.device
    dvec = torch.empty(3, HD, dtype=torch.float32, device=dev)
    dK, dV = torch.empty_like(Kx), torch.empty_like(Vx)
    nbytes = lib.morig_attn_cls_bwd_workspace(N, HD)
    ws = _ws(nbytes, dev)
    dout = dout.contiguous()
    _lib.check(lib.morig_attn_cls_bwd(q0.data_ptr(), kc.data_ptr(), vc.data_ptr(), Kx.data_ptr(), Vx.data_ptr(), att.data_ptr(),
                                      dout.data_ptr(), N, T, HD, d, dvec[0].data_ptr(), dvec[1].data_ptr(),
                                      dvec[2].data_ptr(), dK.data_ptr(), dV.data_ptr(), ws.data_ptr(), nbytes, _sp()),
               "morig_attn_cls_bwd")
    return dvec[0], dvec[1], dvec[2], dK, dV


def concat_cols(xs, pad_to: int = 1) -> torch.Tensor:
    """torch.cat(xs, dim=1) through the strided-copy kernel; with pad_to > 1 the width is rounded up and the extra
    columns are zero (so that the consuming Linear qualifies for the tensor-core engine: K % 4 == 0)"""
    lib = _lib.load()
    xs = [mat(x) for x in xs]
    R = xs[0].shape[0]
    width = sum(x.shape[1] for x in xs)
    total = (width + pad_to - 1) // pad_to * pad_to
    out = (torch.zeros if total != width else torch.empty)(R, total, dtype=torch.float32, device=xs[0].device)
    off = 0
    for x in xs:
        _lib.check(lib.morig_gather_cols(x.data_ptr(), _ld(x), 0, 0, 0, x.shape[1], R, 1, out.data_ptr(), total, off, 0, _sp()),
                   "morig_gather_cols")
        off += x.shape[1]
    return out


def scale_cols(x: torch.Tensor, factor: float) -> torch.Tensor:
    """x * factor through the column-affine kernel of the dense engine's tiny-K path is overkill: one strided copy with a
    per-column scale (morig_bn_train_fwd's apply step exposed through morig_gather_cols would not scale), so this uses
    the BatchNorm-apply entry with scale = factor, shift = 0"""
    x = mat(x)
    R, C = x.shape
    dev = x.device
    sc = torch.full((C,), float(factor), dtype=torch.float32, device=dev)
    sh = torch.zeros(C, dtype=torch.float32, device=dev)
    y = torch.empty(R, C, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().morig_col_affine(x.data_ptr(), _ld(x), R, C, sc.data_ptr(), sh.data_ptr(), y.data_ptr(), C, _sp()),
               "morig_col_affine")
    return y
