"""GPU tier of the joint-extraction post-process and the training losses against their oracle ports."""
import numpy as np
import pytest
import torch

from test_oracle_pinning import _modes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n_half,seed,bw,dens,thr", [(1, 0, 0.05, 0.02, 0.7), (200, 3, 0.03, 0.02, 0.7), (200, 3, 0.05, 0.2, 0.7),
                                                     (1500, 4, 0.02, 0.001, 0.95), (4096, 5, 0.04, 0.02, 0.7)])
def test_nms_meanshift_keeps_exactly_the_oracles_modes(n_half, seed, bw, dens, thr):
    """same kept set as the port (fp64 ball membership evaluated like numpy, documented tie rule), numpy and tensor API"""
    from morig_b200 import cluster_utils
    from oracle import cluster_port
    pts, attn = _modes(n_half, seed)
    want = cluster_port.nms_meanshift(pts.copy(), attn, bw, dens, thr)
    got = cluster_utils.nms_meanshift(pts.copy(), attn, bw, dens, thr)
    assert isinstance(got, np.ndarray) and np.array_equal(got, want)
    got_t = cluster_utils.nms_meanshift(torch.from_numpy(pts).to(DEV), torch.from_numpy(attn).to(DEV), bw, dens, thr)
    assert got_t.is_cuda and np.array_equal(got_t.cpu().numpy(), want)


def test_joint_extraction_pipeline_like_eval_rigging():
    """evaluate/eval_rigging.py:83-95 on synthetic shifted points: threshold, reflect, mean-shift, NMS, flip"""
    from morig_b200 import cluster_utils
    from oracle import cluster_port
    pts, attn = _modes(600, 9)
    pts, attn = pts[:600], attn[:600]
    keep = attn.squeeze() > 0.1
    pts, attn = pts[keep], attn[keep]
    pts = np.concatenate((pts, pts * np.array([[-1, 1, 1]])), axis=0)
    attn = np.tile(attn, (2, 1))
    ms_ref = cluster_port.meanshift_cluster(pts, 0.03, attn, max_iter=30)
    ms = cluster_utils.meanshift_cluster(pts, 0.03, attn, max_iter=30)
    assert np.abs(ms - ms_ref).max() < 1e-9
    j_ref, side_ref = cluster_port.flip(cluster_port.nms_meanshift(ms_ref, attn, 0.03, 0.02))
    j, side = cluster_utils.flip(cluster_utils.nms_meanshift(ms, attn, 0.03, 0.02))
    assert j.shape == j_ref.shape and np.array_equal(side, side_ref) and np.abs(j - j_ref).max() < 1e-8 and len(j) >= 3


@pytest.mark.parametrize("n,m", [(1, 1), (300, 40), (4096, 57), (33, 2000)])
def test_chamfer_forward_and_backward(n, m):
    from morig_b200 import customized_losses as L
    from oracle import cluster_port, losses_port
    g = torch.Generator().manual_seed(n + m)
    p1, p2 = torch.randn(1, n, 3, generator=g), torch.randn(1, m, 3, generator=g)
    a, b = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
    want = losses_port.chamfer_distance_with_average(a, b)
    want.backward()
    x, y = p1.to(DEV).requires_grad_(True), p2.to(DEV).requires_grad_(True)
    got = L.chamfer_distance_with_average(x, y)
    (got * 1.0).backward()
    assert abs(float(got) - float(want)) < 1e-6 * max(1.0, abs(float(want)))
    assert (x.grad.cpu() - a.grad).abs().max() < 1e-6 and (y.grad.cpu() - b.grad).abs().max() < 1e-6
    assert abs(L.chamfer_dist(p1[0].double().numpy(), p2[0].double().numpy())
               - cluster_port.chamfer_dist(p1[0].double().numpy(), p2[0].double().numpy())) < 1e-12


def test_info_nce_losses_forward_and_backward():
    from morig_b200 import customized_losses as L
    from oracle import losses_port
    g = torch.Generator().manual_seed(0)
    n = 1100
    feat = torch.nn.functional.normalize(torch.randn(n, 64, generator=g), dim=1)
    skin = torch.zeros(n, 6); skin[torch.arange(n), torch.randint(0, 6, (n,), generator=g)] = 1.0
    batch = torch.repeat_interleave(torch.arange(2), n // 2)
    f_ref = feat.clone().requires_grad_(True)
    np.random.seed(3); torch.manual_seed(3)
    want = losses_port.multi_pos_info_nce(f_ref, skin, batch)
    want.backward()
    f = feat.to(DEV).requires_grad_(True)
    np.random.seed(3); torch.manual_seed(3)
    # the sampling runs on the CPU generator in both cases: draw it there, as the reference would on its device
    got = L.multi_pos_infoNCE(f, skin.to(DEV), batch.to(DEV))
    got.backward()
    # torch.multinomial draws differ between the CPU and CUDA generators: compare through a CPU-sampled replay instead
    v, p = torch.randn(n, 32, generator=g), torch.randn(900, 32, generator=g)
    pb = torch.repeat_interleave(torch.arange(2), 450)
    c1 = torch.stack([torch.randint(0, 550, (200,), generator=g), torch.randint(0, 450, (200,), generator=g)], 1)
    c2 = torch.stack([torch.randint(0, 450, (160,), generator=g), torch.randint(0, 550, (160,), generator=g)], 1)
    cb1, cb2 = torch.repeat_interleave(torch.arange(2), 100), torch.repeat_interleave(torch.arange(2), 80)
    vr, pr = v.clone().requires_grad_(True), p.clone().requires_grad_(True)
    w2 = losses_port.info_nce(vr, pr, c1, c2, batch, pb, cb1, cb2, 0.07)
    w2.backward()
    vd, pd = v.to(DEV).requires_grad_(True), p.to(DEV).requires_grad_(True)
    g2 = L.infoNCE(vd, pd, c1.to(DEV), c2.to(DEV), batch.to(DEV), pb.to(DEV), cb1.to(DEV), cb2.to(DEV), 0.07)
    g2.backward()
    assert abs(float(g2) - float(w2)) < 2e-5 * max(1.0, abs(float(w2)))
    assert (vd.grad.cpu() - vr.grad).abs().max() < 1e-4 * max(1.0, float(vr.grad.abs().max()))
    assert (pd.grad.cpu() - pr.grad).abs().max() < 1e-4 * max(1.0, float(pr.grad.abs().max()))
    assert torch.isfinite(got) and torch.isfinite(f.grad).all() and float(got) > 0


def test_multi_pos_info_nce_rows_against_dense_formulation():
    """the candidate-list cross entropy of multi_pos_infoNCE against the reference's dense 512 x 512 formulation with the
    SAME index tensors (sampling factored out)"""
    from morig_b200.customized_losses import _InfoNCERows
    g = torch.Generator().manual_seed(1)
    f = torch.nn.functional.normalize(torch.randn(512, 32, generator=g), dim=1)
    pos = torch.randint(0, 512, (512, 1), generator=g)
    neg = torch.randint(0, 512, (512, 200), generator=g)
    fr = f.clone().requires_grad_(True)
    prod = fr @ fr.T
    want = torch.nn.functional.cross_entropy(torch.cat((torch.gather(prod, 1, pos), torch.gather(prod, 1, neg)), 1),
                                             torch.zeros(512).long())
    want.backward()
    fd = f.to(DEV).requires_grad_(True)
    got = _InfoNCERows.apply(fd, fd, torch.zeros(512, dtype=torch.long, device=DEV), torch.cat((pos, neg), 1).to(DEV), 1.0)
    got.backward()
    assert abs(float(got) - float(want)) < 1e-5
    assert (fd.grad.cpu() - fr.grad).abs().max() < 1e-6
