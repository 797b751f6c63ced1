"""Autograd functions of the training path: `torch.autograd.Function`s whose forward and backward
launch the kernels of `libmorig_b200.so` through `train_ops`, so that `loss.backward()` in the reference's training loops
(training/train_rig.py:136-195, training/train_skin.py:139-183) runs on this package's CUDA path.

Granularity follows the reference's blocks:
  LinReluBN      one `MLP` block, Linear -> ReLU -> BatchNorm1d(train)              (models/basic_modules.py:31-36)
  Linear         bare Linear (heads, attention projections)
  EdgeGatherRelu first edge Linear after factorisation: relu(P[i] + Q[j]) per edge  (models/basic_modules.py:193-194)
  BNTrain        train-mode BatchNorm1d on an already rectified input (the per-edge layer above)
  SegMax         max aggregation over CSR segments / graphs, gradient to the FIRST maximal row (PyG aggr='max',
                 torch_scatter.scatter_max: models/basic_modules.py:180-181, models/rignet.py:63,176)
  RowGather      repeat_interleave(x_global, bincount(batch))                       (models/rignet.py:64)
  Normalize      F.normalize(dim=1)                                                 (models/rignet.py:87,98)
  AttnCls        cls-query attention over the key-frames                            (models/rignet.py:36-45)
  ConcatCols     torch.cat(dim=1)
torch itself only differentiates parameter-sized algebra (weight slicing / subtraction of the factorised first edge
Linear) and sums gradients of tensors with several consumers.
"""
from __future__ import annotations

import torch

from . import train_ops as T


class Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return T.linear_fwd(x, w, b, relu=False)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = T.matmul_nn(dy, w) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = T.wgrad(dy, x, ctx.has_bias)
        return dx, dw, db


class LinReluBN(torch.autograd.Function):
    """y = BatchNorm_train(relu(x w^T + b)); running statistics updated in place (momentum as the module's)"""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, momentum):
        r = T.linear_fwd(x, w, b, relu=True)
        y, mean, invstd = T.bn_train_fwd(r, gamma, beta, running_mean, running_var, momentum)
        ctx.save_for_backward(x, w, r, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, r, gamma, mean, invstd = ctx.saved_tensors
        dz, dgamma, dbeta = T.bn_relu_bwd(dy, r, gamma, mean, invstd, relu=True)
        dx = T.matmul_nn(dz, w) if ctx.needs_input_grad[0] else None
        dw, db = T.wgrad(dz, x, True)
        return dx, dw, db, dgamma, dbeta, None, None, None


class BNTrain(torch.autograd.Function):
    """train-mode BatchNorm1d on x (no ReLU mask in the backward: x is differentiated as is)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum):
        y, mean, invstd = T.bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum)
        ctx.save_for_backward(x, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, invstd = ctx.saved_tensors
        dx, dgamma, dbeta = T.bn_relu_bwd(dy, x, gamma, mean, invstd, relu=False)
        return dx, dgamma, dbeta, None, None, None


class EdgeGatherRelu(torch.autograd.Function):
    """h[e] = relu(PQ[tgt[e], :H] + PQ[col[e], H:]) for every CSR slot of `graph`"""

    @staticmethod
    def forward(ctx, pq, graph):
        H = pq.shape[1] // 2
        h = T.edge_gather_relu(pq[:, :H], pq[:, H:], graph)
        ctx.save_for_backward(h)
        ctx.graph = graph
        return h

    @staticmethod
    def backward(ctx, dh):
        (h,) = ctx.saved_tensors
        return T.edge_gather_relu_bwd(dh, h, ctx.graph), None


class SegMax(torch.autograd.Function):
    """out[s] = max over rows ptr[s]..ptr[s+1]; the gradient goes to the first maximal row of each (segment, column)"""

    @staticmethod
    def forward(ctx, y, ptr, n_seg):
        out, arg = T.segmax_fwd(y, ptr, n_seg)
        ctx.save_for_backward(arg)
        ctx.rows = y.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, dout, _darg):
        (arg,) = ctx.saved_tensors
        return T.segmax_bwd(dout, arg, ctx.rows), None, None


class RowGather(torch.autograd.Function):
    """out[r] = src[idx[r]] for a sorted idx whose segments are given by ptr"""

    @staticmethod
    def forward(ctx, src, idx32, ptr):
        ctx.save_for_backward(ptr)
        ctx.n_seg = src.shape[0]
        return T.row_gather(src, idx32)

    @staticmethod
    def backward(ctx, dout):
        (ptr,) = ctx.saved_tensors
        return T.seg_sum(dout, ptr, ctx.n_seg), None, None


class Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return T.normalize_fwd(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return T.normalize_bwd(x, dy)


class AttnCls(torch.autograd.Function):
    """q0 / kc / vc [1, HD] (projected cls token), Kx / Vx [N, T, HD] (projected key-frame tokens) -> [N, HD]"""

    @staticmethod
    def forward(ctx, q0, kc, vc, Kx, Vx, d):
        q0, kc, vc, Kx, Vx = (t.contiguous() for t in (q0, kc, vc, Kx, Vx))
        out, att = T.attn_cls_fwd(q0, kc, vc, Kx, Vx, d)
        ctx.save_for_backward(q0, kc, vc, Kx, Vx, att)
        ctx.d = d
        return out

    @staticmethod
    def backward(ctx, dout):
        q0, kc, vc, Kx, Vx, att = ctx.saved_tensors
        dq, dk, dv, dK, dV = T.attn_cls_bwd(q0, kc, vc, Kx, Vx, att, dout, ctx.d)
        return dq.view_as(q0), dk.view_as(kc), dv.view_as(vc), dK, dV, None


class ConcatColsPadded(torch.autograd.Function):
    """torch.cat(xs, dim=1) with the width rounded up to a multiple of 4 by zero columns (the consumer pads its weight with
    zero columns accordingly, train_forward.mlp_block), so that wide vertex layers such as mlp_transform.0 (K = 1862 / 1923)
    qualify for the tensor-core engine"""

    @staticmethod
    def forward(ctx, *xs):
        ctx.widths = [x.shape[1] for x in xs]
        return T.concat_cols(xs, 4)

    @staticmethod
    def backward(ctx, dout):
        outs, off = [], 0
        for i, w in enumerate(ctx.widths):
            outs.append(dout[:, off:off + w] if ctx.needs_input_grad[i] else None)
            off += w
        return tuple(outs)


class ConcatCols(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        ctx.widths = [x.shape[1] for x in xs]
        return T.concat_cols(xs)

    @staticmethod
    def backward(ctx, dout):
        outs, off = [], 0
        for i, w in enumerate(ctx.widths):
            outs.append(dout[:, off:off + w] if ctx.needs_input_grad[i] else None)     # views: consumers take row strides
            off += w
        return tuple(outs)


class FrameMean(torch.autograd.Function):
    """mean over the key-frame axis of x [N, T, C] (aggr_method='mean', models/rignet.py:92-93): per-vertex segment sums of
    the [N * T, C] view, scaled by 1 / T (a power-of-two-free constant: applied through the BatchNorm-affine kernel)"""

    @staticmethod
    def forward(ctx, x):
        n, t, c = x.shape
        ctx.shape = (n, t, c)
        ptr = torch.arange(0, n * t + 1, t, dtype=torch.int32, device=x.device)
        s = T.seg_sum(x.reshape(n * t, c), ptr, n)
        return T.scale_cols(s, 1.0 / t)

    @staticmethod
    def backward(ctx, dout):
        n, t, c = ctx.shape
        idx = torch.arange(n, dtype=torch.int32, device=dout.device).repeat_interleave(t)
        return T.row_gather(T.scale_cols(dout, 1.0 / t), idx).view(n, t, c)


class FrameMax(torch.autograd.Function):
    """max over the key-frame axis of x [N, T, C] (aggr_method='max', models/rignet.py:94-95; torch.max(dim=1) sends the
    gradient to the first maximal key-frame)"""

    @staticmethod
    def forward(ctx, x):
        n, t, c = x.shape
        ctx.shape = (n, t, c)
        ptr = torch.arange(0, n * t + 1, t, dtype=torch.int32, device=x.device)
        out, arg = T.segmax_fwd(x.reshape(n * t, c), ptr, n)
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        n, t, c = ctx.shape
        return T.segmax_bwd(dout, arg, n * t).view(n, t, c)
