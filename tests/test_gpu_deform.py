"""GPU tier of the upstream flow producer: point-cloud primitives against their restatements
(oracle/pointops_port.py, bit-exact index work), CorrNet / DeformNet against the fixture produced by the reference's own
unmodified modules (oracle/gen_golden_deform.py), and the surface-sampling front-end of the geodesic build (8(f) #2)."""
import os

import numpy as np
import pytest
import torch

import helpers
from morig_b200 import synth
from oracle import pointops_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cloud(b, n, seed):
    g = torch.Generator().manual_seed(seed)
    sizes = [n + 7 * i for i in range(b)]
    pos = torch.cat([torch.rand(s, 3, generator=g) - 0.5 for s in sizes])
    batch = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)])
    return pos, batch


@pytest.mark.parametrize("b,n,ratio", [(1, 10, 0.5), (3, 300, 0.25), (2, 2048, 0.5), (1, 5000, 0.1)])
def test_fps_bit_exact(b, n, ratio):
    from morig_b200 import pointnet2
    pos, batch = _cloud(b, n, n)
    want = pointops_port.fps(pos, batch, ratio, random_start=False)
    got = pointnet2.fps(pos.to(DEV), batch.to(DEV), ratio, random_start=False)
    assert torch.equal(got.cpu(), want)
    torch.manual_seed(5)
    want = pointops_port.fps(pos, batch, ratio, random_start=True)
    torch.manual_seed(5)
    got = pointnet2.fps(pos.to(DEV), batch.to(DEV), ratio, random_start=True)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("k,cosine,d", [(1, True, 64), (5, True, 64), (3, False, 3), (8, False, 3)])
def test_knn_bit_exact(k, cosine, d):
    from morig_b200 import pointnet2
    g = torch.Generator().manual_seed(k + d)
    x, bx = torch.randn(900, d, generator=g), torch.repeat_interleave(torch.arange(3), 300)
    y, by = torch.randn(700, d, generator=g), torch.sort(torch.randint(0, 3, (700,), generator=g)).values
    want = pointops_port.knn(x, y, k, bx, by, cosine=cosine)
    got = pointnet2.knn(x.to(DEV), y.to(DEV), k, bx.to(DEV), by.to(DEV), cosine=cosine)
    same = (got.cpu() == want).all(0)
    if cosine:          # similarities are fp32 dot products: allow swaps between neighbours that tie to ~1e-6
        xn, yn = torch.nn.functional.normalize(x, dim=1), torch.nn.functional.normalize(y, dim=1)
        a = (yn[want[0]] * xn[want[1]]).sum(1)
        c = (yn[got.cpu()[0]] * xn[got.cpu()[1]]).sum(1)
        assert float((a - c).abs().max()) < 1e-5 and float(same.float().mean()) > 0.995
    else:
        assert bool(same.all())


def test_ball_query_and_interpolation():
    from morig_b200 import _lib, pointnet2
    pos, batch = _cloud(2, 600, 3)
    idx = pointops_port.fps(pos, batch, 0.25, random_start=False)
    want = pointops_port.radius(pos, pos[idx], 0.3, batch, batch[idx], max_num_neighbors=16)
    lib = _lib.load()
    p, c = pos.to(DEV), pos[idx].contiguous().to(DEV)
    ptr = pointnet2.batch_ptr(batch.to(DEV))
    m = idx.numel()
    nbr = torch.empty(m, 16, dtype=torch.int32, device=DEV)
    cnt = torch.empty(m, dtype=torch.int32, device=DEV)
    _lib.check(lib.morig_ball_query(p.data_ptr(), ptr.data_ptr(), c.data_ptr(), batch[idx].to(torch.int32).to(DEV).data_ptr(), m, 0.3, 16,
                                    nbr.data_ptr(), cnt.data_ptr(), _lib.stream_ptr()), "ball_query")
    rows = torch.arange(m).repeat_interleave(cnt.cpu().long())
    cols = torch.cat([nbr.cpu()[i, : int(cnt[i])] for i in range(m)]).long()
    assert torch.equal(torch.stack([rows, cols]), want) and int(cnt.max()) == 16
    f = torch.randn(pos.shape[0], 40, generator=torch.Generator().manual_seed(0))
    want = pointops_port.knn_interpolate(f[idx], pos[idx], pos, batch[idx], batch, k=3)
    got = pointnet2.knn_interpolate(f[idx].contiguous().to(DEV), c, p, batch[idx].to(DEV), batch.to(DEV), k=3)
    assert helpers.max_abs_diff(got, want) < 1e-4 * max(1.0, float(want.abs().max()))


def _fixture(name):
    z = np.load(os.path.join(helpers.GOLDEN_DIR, name + ".npz"))
    data = synth.Batch(vtx=torch.from_numpy(z["vtx"]), pts=torch.from_numpy(z["pts"]),
                       vtx_batch=torch.from_numpy(z["vtx_batch"]).long(), pts_batch=torch.from_numpy(z["pts_batch"]).long(),
                       tpl_edge_index=torch.from_numpy(z["tpl_edge_index"]).long(),
                       geo_edge_index=torch.from_numpy(z["geo_edge_index"]).long(), num_graphs=int(z["graphs"]))
    return z, data


def test_deformnet_matches_fixture_of_the_unmodified_reference():
    """CorrNet (GCU chain + PointNet++ branch + visibility head) and DeformNet (cosine kNN interpolation + GCNDeform)
    against outputs of the reference's own modules (its GPU code branch, third-party operators restated)"""
    import morig_b200
    z, data = _fixture("deformnet_b2_v400_p300")
    chk = synth.make_deform_batch(int(z["graphs"]), int(z["n_vtx"]), int(z["n_pts"]), seed=int(z["data_seed"]))
    assert torch.equal(chk.vtx, data.vtx) and torch.equal(chk.pts, data.pts)          # generator is pinned
    net = morig_b200.deformnet(tau_nce=0.07, num_interp=5).eval()       # the factory, as models.__dict__["deformnet"]
    net.load_state_dict(synth.seeded_state_dict(net, int(z["weight_seed"])))
    net = net.to(DEV)
    torch.manual_seed(int(z["rng_seed"]))
    with torch.no_grad():
        pred_flow, vtx_f, pts_f, vis, tau = net(data.to(DEV))
    assert helpers.max_abs_diff(vtx_f, torch.from_numpy(z["vtx_feature"])) < 1e-4
    assert helpers.max_abs_diff(pts_f, torch.from_numpy(z["pts_feature"])) < 1e-4
    # the visibility head sees the 1-nearest point FEATURE of every vertex (a hard selection among near-ties) and is
    # min-max normalised per sample afterwards, which stretches small differences
    assert helpers.max_abs_diff(vis, torch.from_numpy(z["pred_vismask"])) < 2e-2
    assert float((vis.cpu() - torch.from_numpy(z["pred_vismask"])).abs().mean()) < 2e-3
    # pred_flow depends on hard thresholds (visible / invisible at 0.5) and top-5 selections: compare where the
    # visibility decision is not marginal
    ref_vis = torch.from_numpy(z["pred_vismask"]).squeeze(1)
    solid = (ref_vis - 0.5).abs() > 0.02
    err = (pred_flow.cpu() - torch.from_numpy(z["pred_flow"])).abs().max(1).values
    print("pred_flow: max err (solid)", float(err[solid].max()), "mean err", float(err.mean()), "frac < 2e-3",
          float((err < 2e-3).float().mean()))
    assert float(err[solid].mean()) < 1e-3 and float((err < 5e-3).float().mean()) > 0.97
    assert float(tau) == pytest.approx(0.07)


def test_surface_sampling_front_end():
    """8(f) #2 front-end: area-weighted surface samples + normals, thinned by farthest point sampling (the stand-in for
    open3d's Poisson-disk sampler, data_proc/common_ops.py:175-181): samples lie on their triangles, normals are unit
    face normals, the thinned set is well spread, everything is reproducible; feeds calc_surface_geodesic end to end"""
    from morig_b200 import graph_build
    nu, nv = 24, 20
    verts = synth.torus_vertices(nu * nv, np.random.default_rng(0)).astype(np.float64)
    iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a, b = (iu * nv + iv).ravel(), (((iu + 1) % nu) * nv + iv).ravel()
    c, d = (iu * nv + (iv + 1) % nv).ravel(), (((iu + 1) % nu) * nv + (iv + 1) % nv).ravel()
    faces = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)]).astype(np.int64)
    pts, nrm = graph_build.sample_surface_poisson(verts, faces, 600, seed=3)
    pts2, nrm2 = graph_build.sample_surface_poisson(verts, faces, 600, seed=3)
    assert pts.shape == (600, 3) and np.array_equal(pts, pts2) and np.array_equal(nrm, nrm2)
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-12
    # every sample lies in the plane of (and inside) some triangle: distance to the vertex set is below the longest edge
    edge = max(np.linalg.norm(verts[faces[:, i]] - verts[faces[:, (i + 1) % 3]], axis=1).max() for i in range(3))
    d = np.sqrt(((pts[:, None] - verts[None]) ** 2).sum(-1)).min(1)
    assert d.max() <= edge
    # blue-noise property: the minimum pairwise distance of the thinned set beats plain uniform sampling by a wide margin
    raw, _ = graph_build.sample_surface_uniform(verts, faces, 600, seed=3)
    def min_dist(p):
        dd = np.sqrt(((p[:, None] - p[None]) ** 2).sum(-1)); np.fill_diagonal(dd, 1e9); return dd.min()
    assert min_dist(pts) > 5 * min_dist(raw)
    geo = graph_build.calc_surface_geodesic(verts, faces, number_of_points=600, seed=3)
    assert geo.shape == (len(verts), len(verts)) and np.array_equal(geo, geo.T) and np.isfinite(geo).all()
    edges = graph_build.get_geo_edges(geo, verts, seed=0)
    assert edges.shape[1] == 2 and np.bincount(edges[:, 0], minlength=len(verts)).max() <= 15


@pytest.mark.parametrize("radius,max_nn", [(0.06, 15), (0.3, 4), (100.0, 3)])
def test_get_geo_edges_reproduces_the_seeded_random_subset(radius, max_nn):
    """reference rule of data_proc/common_ops.py:220-222 (np.random.choice in vertex order): same global-generator state
    -> the same edge list as the oracle port (which is pinned to the unmodified reference function)"""
    from morig_b200 import graph_build
    from oracle import gen_golden_geodesic as gg
    from oracle import geodesic_port
    pts, nrm, verts = gg.make_inputs(600, 300, 4)
    geo = geodesic_port.surface_geodesic_from_samples(pts, nrm, verts)
    np.random.seed(7)
    want = geodesic_port.geo_edges_random_subset(geo, radius, max_nn)
    np.random.seed(7)
    got = graph_build.get_geo_edges(geo, verts, radius, max_nn)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    assert np.array_equal(graph_build.get_geo_edges(geo, verts, radius, max_nn, seed=7), want)
