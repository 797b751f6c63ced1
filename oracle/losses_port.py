"""TEST INFRASTRUCTURE (oracle): torch-CPU restatements of the two training losses of the rigging networks,
models/customized_losses.py:107-158,231-251, pinned in the build container against the unmodified reference functions
(tests/test_oracle_pinning.py)."""
import numpy as np
import torch


def chamfer_distance_with_average(p1, p2):
    """models/customized_losses.py:231-251"""
    assert p1.size(0) == 1 and p2.size(0) == 1 and p1.size(2) == p2.size(2)
    a = p1.repeat(p2.size(1), 1, 1).transpose(0, 1)                  # [N, M, D]
    b = p2.repeat(a.size(0), 1, 1)                                   # [N, M, D]
    dist_norm = torch.norm(a - b, 2, dim=2)
    return 0.5 * (torch.mean(torch.min(dist_norm, dim=1)[0]) + torch.mean(torch.min(dist_norm, dim=0)[0]))


def info_nce(vtx_feature, pts_feature, corr_v2p, corr_p2v, vtx_batch, pts_batch, corr_v2p_batch, corr_p2v_batch, tau):
    """models/customized_losses.py:107-135"""
    ce = torch.nn.CrossEntropyLoss(reduction="none")
    loss = 0.0
    for i in range(len(torch.unique(vtx_batch))):
        v, p = vtx_feature[vtx_batch == i], pts_feature[pts_batch == i]
        c = corr_v2p[corr_v2p_batch == i]
        if len(c) == 0:
            continue
        loss = loss + ce(torch.mm(v[c[:, 0]], p.T) / tau, c[:, 1]).mean()
        c = corr_p2v[corr_p2v_batch == i]
        if len(c) == 0:
            continue
        loss = loss + ce(torch.mm(p[c[:, 0]], v.T) / tau, c[:, 1]).mean()
    return loss / len(torch.unique(vtx_batch))


def multi_pos_info_nce(pred_feature, gt_skin, batch):
    """models/customized_losses.py:137-158 (same RNG calls in the same order)"""
    ce = torch.nn.CrossEntropyLoss(reduction="mean")
    loss = 0.0
    for i in range(len(torch.unique(batch))):
        ids = np.random.choice((batch == i).sum().item(), 512, replace=False)
        f = pred_feature[batch == i][ids]
        s = gt_skin[batch == i][ids]
        sim = ((2 - torch.sum(torch.abs(s[None] - s[:, None]), axis=-1)) / 2.0 > 0.9).float()
        pos = torch.multinomial(sim, 10, replacement=True)
        neg = torch.multinomial(1 - sim, 200, replacement=True)
        prod = torch.mm(f, f.T)
        prod_neg = torch.gather(prod, dim=1, index=neg)
        loss_i = 0.0
        for j in range(10):
            prod_pos = torch.gather(prod, dim=1, index=pos[:, j][:, None])
            loss_i = loss_i + ce(torch.cat((prod_pos, prod_neg), dim=1), torch.zeros(512).long())
        loss = loss + loss_i / 10
    return loss / len(torch.unique(batch))
