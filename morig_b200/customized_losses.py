"""The two losses of the rigging-network training loops on the GPU path, with the
reference's names and signatures (models/customized_losses.py):

    chamfer_distance_with_average(p1, p2)          :231-251   joint loss of training/train_rig.py:177
    infoNCE(vtx_feature, pts_feature, corr_v2p, corr_p2v, vtx_batch, pts_batch, corr_v2p_batch, corr_p2v_batch, tau)  :107-135
    multi_pos_infoNCE(pred_feature, gt_skin, batch)  :137-158  embedding loss of training/train_rig.py:172-174

plus the numpy `chamfer_dist(pts1, pts2)` of the evaluation (utils/mst_utils.py:316-321).  The reference materialises the
[N, M, D] difference tensor / the [R, M] logit matrix; here nearest neighbours and the log-sum-exp are streamed
(csrc/postproc.cu) and the gradients come from hand-written backward kernels.  The random sampling of
`multi_pos_infoNCE` (np.random.choice + torch.multinomial) is kept call for call, so that seeded runs draw the same
samples as the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _nn_dist(a: torch.Tensor, b: torch.Tensor):
    """nearest-neighbour distance and index of every row of a [N, D] among b [M, D] (fp32 or fp64 CUDA tensors)"""
    lib = _lib.load()
    n, d = a.shape
    m = b.shape[0]
    dist = torch.empty(n, dtype=a.dtype, device=a.device)
    arg = torch.empty(n, dtype=torch.int32, device=a.device)
    fn = lib.morig_nn_dist_f32 if a.dtype == torch.float32 else lib.morig_nn_dist_f64
    with torch.cuda.device(a.device):
        _lib.check(fn(a.data_ptr(), n, b.data_ptr(), m, d, dist.data_ptr(), arg.data_ptr(), _lib.stream_ptr()), "morig_nn_dist")
    return dist, arg


class _ChamferAverage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2):
        a, b = p1.contiguous(), p2.contiguous()
        d1, n1 = _nn_dist(a, b)
        d2, n2 = _nn_dist(b, a)
        ctx.save_for_backward(a, b, d1, n1, d2, n2)
        return _mean(d1, d2)

    @staticmethod
    def backward(ctx, g):
        a, b, d1, n1, d2, n2 = ctx.saved_tensors
        lib = _lib.load()
        n, d = a.shape
        m = b.shape[0]
        gv = float(g)                                                 # scalar upstream gradient
        grads = [None, None]
        with torch.cuda.device(a.device):
            for idx, (x, y, dx, nx, dy, ny, cx, cy) in enumerate(((a, b, d1, n1, d2, n2, n, m), (b, a, d2, n2, d1, n1, m, n))):
                if not ctx.needs_input_grad[idx]:
                    continue
                out = torch.empty_like(x)
                _lib.check(lib.morig_chamfer_bwd_f32(x.data_ptr(), cx, y.data_ptr(), cy, d, dx.data_ptr(), nx.data_ptr(),
                                                     dy.data_ptr(), ny.data_ptr(), 0.5 * gv / cx, 0.5 * gv / cy, out.data_ptr(),
                                                     _lib.stream_ptr()), "morig_chamfer_bwd_f32")
                grads[idx] = out
        return tuple(grads)


def _mean(d1: torch.Tensor, d2: torch.Tensor) -> torch.Tensor:
    # two tiny reductions over already reduced vectors
    return 0.5 * (d1.mean() + d2.mean())


def chamfer_distance_with_average(p1: torch.Tensor, p2: torch.Tensor) -> torch.Tensor:
    """`chamfer_distance_with_average(p1 [1, N, D], p2 [1, M, D])` -- models/customized_losses.py:231-251:
    0.5 * (mean_i min_j |p1_i - p2_j| + mean_j min_i |p1_i - p2_j|), differentiable in both arguments."""
    assert p1.size(0) == 1 and p2.size(0) == 1 and p1.size(2) == p2.size(2)
    if not p1.is_cuda:
        raise RuntimeError("morig_b200.customized_losses: CUDA tensors expected (no CPU path)")
    return _ChamferAverage.apply(p1[0].float(), p2[0].float())


def chamfer_dist(pts1, pts2):
    """`chamfer_dist(pts1, pts2)` -- utils/mst_utils.py:316-321 (fp64 numpy in the reference's evaluation)"""
    as_numpy = isinstance(pts1, np.ndarray)
    if as_numpy:
        if not torch.cuda.is_available():
            raise RuntimeError("morig_b200: a CUDA device is required (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
        a = torch.from_numpy(np.ascontiguousarray(pts1, dtype=np.float64)).to(dev)
        b = torch.from_numpy(np.ascontiguousarray(pts2, dtype=np.float64)).to(dev)
    else:
        a, b = pts1.double().contiguous(), pts2.double().contiguous()
    d1, _ = _nn_dist(a, b)
    d2, _ = _nn_dist(b, a)
    out = 0.5 * (d1.mean() + d2.mean())
    return float(out) if as_numpy else out


class _InfoNCERows(torch.autograd.Function):
    """mean over rows of  logsumexp_m(A[r] . K[m] / tau) - A[r] . K[label[r]] / tau  (candidate lists optional)"""

    @staticmethod
    def forward(ctx, A, K, label, sel, tau):
        lib = _lib.load()
        A, K = A.contiguous(), K.contiguous()
        label = label.contiguous()
        R, C = A.shape
        M = K.shape[0]
        S = 0 if sel is None else sel.shape[1]
        sel = None if sel is None else sel.contiguous()
        loss = torch.empty(R, dtype=torch.float32, device=A.device)
        lse = torch.empty(R, dtype=torch.float32, device=A.device)
        with torch.cuda.device(A.device):
            _lib.check(lib.morig_info_nce_fwd(A.data_ptr(), C, K.data_ptr(), K.shape[1], label.data_ptr(), _lib.ptr(sel), S, R, M, C,
                                              float(tau), loss.data_ptr(), lse.data_ptr(), _lib.stream_ptr()), "morig_info_nce_fwd")
        ctx.save_for_backward(A, K, label, lse)
        ctx.sel, ctx.tau = sel, float(tau)
        return loss.mean()

    @staticmethod
    def backward(ctx, g):
        A, K, label, lse = ctx.saved_tensors
        lib = _lib.load()
        R, C = A.shape
        M = K.shape[0]
        sel = ctx.sel
        S = 0 if sel is None else sel.shape[1]
        gr = torch.full((R,), float(g) / R, dtype=torch.float32, device=A.device)
        dA = torch.empty_like(A)
        dK = torch.empty_like(K) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(A.device):
            _lib.check(lib.morig_info_nce_bwd(A.data_ptr(), C, K.data_ptr(), K.shape[1], label.data_ptr(), _lib.ptr(sel), S,
                                              lse.data_ptr(), gr.data_ptr(), R, M, C, ctx.tau, dA.data_ptr(), C, _lib.ptr(dK),
                                              K.shape[1], _lib.stream_ptr()), "morig_info_nce_bwd")
        return dA, dK, None, None, None


def infoNCE(vtx_feature, pts_feature, corr_v2p, corr_p2v, vtx_batch, pts_batch, corr_v2p_batch, corr_p2v_batch, tau):
    """`infoNCE(...)` -- models/customized_losses.py:107-135 (correspondence loss of training/train_corr*.py): per sample,
    cross entropy of anchor . candidates^T / tau in both directions."""
    loss = 0.0
    for i in range(len(torch.unique(vtx_batch))):
        v = vtx_feature[vtx_batch == i]
        p = pts_feature[pts_batch == i]
        c_v2p = corr_v2p[corr_v2p_batch == i]
        if len(c_v2p) == 0:
            loss += 0.0
            continue
        loss = loss + _InfoNCERows.apply(v[c_v2p[:, 0]], p, c_v2p[:, 1], None, tau)
        c_p2v = corr_p2v[corr_p2v_batch == i]
        if len(c_p2v) == 0:
            loss += 0.0
            continue
        loss = loss + _InfoNCERows.apply(p[c_p2v[:, 0]], v, c_p2v[:, 1], None, tau)
    return loss / len(torch.unique(vtx_batch))


def multi_pos_infoNCE(pred_feature, gt_skin, batch):
    """`multi_pos_infoNCE(pred_feature, gt_skin, batch)` -- models/customized_losses.py:137-158.  Sampling (512 vertices per
    mesh, 10 positives and 200 negatives per anchor) is drawn exactly like the reference; each of the 10 cross
    entropies runs over the anchor's candidate list [positive_j | negatives] without materialising the 512 x 512
    product."""
    loss = 0.0
    for i in range(len(torch.unique(batch))):
        sample_ids = np.random.choice((batch == i).sum().item(), 512, replace=False)
        feature_i = pred_feature[batch == i][sample_ids]
        gt_skin_i = gt_skin[batch == i][sample_ids]
        gt_sim = (2 - torch.sum(torch.abs(gt_skin_i[None] - gt_skin_i[:, None]), axis=-1)) / 2.0
        gt_sim = (gt_sim > 0.9).float()
        pos_ids = torch.multinomial(gt_sim, 10, replacement=True)
        neg_ids = torch.multinomial(1 - gt_sim, 200, replacement=True)
        zeros = torch.zeros(512, dtype=torch.long, device=pred_feature.device)
        loss_i = 0.0
        for j in range(10):
            sel = torch.cat((pos_ids[:, j][:, None], neg_ids), dim=1)
            loss_i = loss_i + _InfoNCERows.apply(feature_i, feature_i, zeros, sel, 1.0)
        loss = loss + loss_i / 10
    return loss / len(torch.unique(batch))
