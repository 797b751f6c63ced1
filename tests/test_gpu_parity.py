"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle and the golden
fixtures produced by the unmodified reference.  Run with `pytest -m gpu` on a B200."""
import os

import numpy as np
import pytest
import torch

import helpers
from morig_b200 import _lib, engine, packing, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_library_loaded_is_in_tree():
    lib = _lib.load()
    assert lib.morig_version() == _lib.ABI_VERSION
    assert lib.morig_sm_count() > 0


# ---- integer work: bit-exact ---------------------------------------------------------------------
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (7, 30, 1), (64, 500, 2), (1000, 9000, 3), (4096, 61440, 4)])
def test_graph_prep_bit_exact(n, e, seed):
    from oracle import graph_port
    rng = np.random.default_rng(seed)
    ei = rng.integers(0, n, size=(2, e)).astype(np.int64)          # duplicates and self loops included
    if e:
        ei[1, : e // 4] = ei[1, 0]                                   # one heavy target
    g = engine.graph_prep(torch.from_numpy(ei).to(DEV), n)
    rowptr, col = graph_port.csr_by_target(ei, n)
    e_real = int(rowptr[-1])
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.col.cpu().numpy()[:e_real], col)
    assert np.array_equal(g.tgt.cpu().numpy()[:e_real], np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr)))


def test_graph_prep_hub_with_64k_in_edges():
    """degree robustness: a star whose hub has 65 536 in-edges (plus duplicates and self loops) -- the CSR is a stable
    radix sort by target, so the cost does not depend on the in-degree distribution; bit-exact vs the oracle"""
    from oracle import graph_port
    n = 70000
    rng = np.random.default_rng(7)
    src = rng.permutation(n)[:65536]
    ei = np.stack([src, np.full(65536, 5)]).astype(np.int64)
    extra = rng.integers(0, n, size=(2, 20000)).astype(np.int64)
    ei = np.concatenate([ei[:, :30000], extra, ei[:, 30000:], ei[:, :100]], axis=1)
    g = engine.graph_prep(torch.from_numpy(ei).to(DEV), n)
    rowptr, col = graph_port.csr_by_target(ei, n)
    e_real = int(rowptr[-1])
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.col.cpu().numpy()[:e_real], col)
    assert np.array_equal(g.tgt.cpu().numpy()[:e_real], np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr)))


def test_graph_prep_on_dataset_style_input():
    """edge lists as the dataset delivers them (self loops already appended, datasets/dataset_rig.py:121-122)"""
    from oracle import graph_port
    m = synth.make_mesh(1024, 0)
    for key in ("tpl_edge_index", "geo_edge_index"):
        ei = m[key]
        g = engine.graph_prep(torch.from_numpy(ei).to(DEV), 1024)
        rowptr, col = graph_port.csr_by_target(ei, 1024)
        assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
        assert np.array_equal(g.col.cpu().numpy()[: rowptr[-1]], col)
        assert int(rowptr[-1]) == ei.shape[1]


@pytest.mark.parametrize("n,b", [(256, 1), (1024, 2), (4096, 1)])
def test_knn_graph_bit_exact(n, b):
    lib = _lib.load()
    pos = np.concatenate([synth.torus_vertices(n, np.random.default_rng(s)) for s in range(b)])
    gptr = torch.arange(0, (b + 1) * n, n, dtype=torch.int32, device=DEV)
    out = torch.empty(2, 15 * n * b, dtype=torch.int64, device=DEV)
    p = torch.from_numpy(pos).to(DEV)
    _lib.check(lib.morig_knn_graph(p.data_ptr(), gptr.data_ptr(), b, n * b, 15, out.data_ptr(), _lib.stream_ptr()), "knn")
    ref = np.concatenate([synth.knn_edges(pos[s * n:(s + 1) * n], 15) + s * n for s in range(b)], axis=1)
    assert np.array_equal(out.cpu().numpy(), ref)


# ---- kernels against straightforward torch fp32 references -------------------------------------------
@pytest.mark.parametrize("M,K,N", [(1, 3, 5), (130, 36, 64), (257, 838, 1024), (1000, 256, 3), (4099, 544, 512),
                                   (20, 1024, 1024), (4, 1024, 1024), (32, 64, 9), (1, 1024, 1021),    # skinny kernel
                                   (1000, 3, 192), (257, 4, 128), (5, 8, 12), (4099, 3, 64)])          # tiny-K kernel
def test_dense_fwd(M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b, s, t = torch.randn(N, generator=g), torch.randn(N, generator=g), torch.randn(N, generator=g)
    layer = packing.DenseLayer(W=packing._pack_wt(W.double()).to(DEV), K=K, N=N, bias=b.to(DEV), scale=s.to(DEV),
                               shift=t.to(DEV), relu=True)
    C = torch.empty(M, N, device=DEV)
    engine.dense(layer, A.to(DEV), 0, K, M, C=C, ldc=N)
    ref = torch.relu(A.double() @ W.double().t() + b.double()) * s.double() + t.double()
    assert helpers.max_abs_diff(C, ref) < 2e-5


KINDS = [pytest.param(packing.KIND_TF32, id="tf32"), pytest.param(packing.KIND_F16, id="f16")]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("M,K,N", [(128, 32, 64), (300, 96, 64), (1000, 64, 128), (257, 288, 256), (4099, 544, 512),
                                   (640, 840, 1024), (20, 1024, 1024), (513, 36, 768), (40000, 256, 256),
                                   (300, 64, 40), (500, 128, 100), (700, 96, 200), (129, 32, 17), (33, 64, 1000)])
def test_dense_fwd_tensor_core(M, K, N, kind):
    """tcgen05 split-precision engine (both operand kinds) against an fp64 reference: fp32-class accuracy is
    required (tolerance of the path is 1e-4 absolute after ~20 chained layers, so a single layer must stay near
    fp32 rounding)"""
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b, s, t = torch.randn(N, generator=g), torch.randn(N, generator=g), torch.randn(N, generator=g)
    layer = packing.DenseLayer(W=packing._pack_wt(W.double()).to(DEV), K=K, N=N, bias=b.to(DEV), scale=s.to(DEV),
                               shift=t.to(DEV), relu=True).with_tc(W.double(), kind)
    assert layer.Wtc is not None and layer.tc_kind == kind
    layer.Wtc = layer.Wtc.to(DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    engine.dense(layer, A.to(DEV), 0, K, M, C=C, ldc=N)
    ref = torch.relu(A.double() @ W.double().t() + b.double()) * s.double() + t.double()
    # 3xTF32 keeps ~21 mantissa bits per product but the tensor core accumulates K/8 partial sums with
    # truncation: allow 1e-5 of the output range (plain TF32 would be ~1e-3, fp32 FFMA ~1e-6)
    assert helpers.max_abs_diff(C, ref) < 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("a_mag,w_mag", [(3.0e4, 1.0), (1.0e-6, 1.0), (1.0, 2.0e3), (250.0, 1.0e-4), (1.0e6, 1.0e-5)])
def test_dense_fp16_kind_dynamic_range(a_mag, w_mag):
    """fp16 operand kind: activations and weights far outside fp16's exponent range must be handled by the
    power-of-two operand scales (derived on the device from the tracked max |A|), and columns whose magnitudes
    differ by 2^12 inside one tensor must all keep fp32-class accuracy"""
    M, K, N = 1500, 320, 192
    g = torch.Generator().manual_seed(5)
    col_scale = torch.ones(K)
    col_scale[::3] = 2.0 ** -12
    A = torch.randn(M, K, generator=g) * col_scale * a_mag
    W = torch.randn(N, K, generator=g) / K ** 0.5 * w_mag
    W = W / col_scale.clamp(min=2.0 ** -6)                       # small activation columns meet larger weights
    layer = packing.DenseLayer(W=packing._pack_wt(W.double()).to(DEV), K=K, N=N).with_tc(W.double(), packing.KIND_F16)
    layer.Wtc = layer.Wtc.to(DEV)
    C = torch.full((M, N), float("nan"), device=DEV)
    engine.dense(layer, A.to(DEV), 0, K, M, C=C, ldc=N)
    ref = A.double() @ W.double().t()
    assert torch.isfinite(C).all()
    assert helpers.max_abs_diff(C, ref) < 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("H,frames,mag", [(64, 1, 1.0), (128, 2, 300.0), (256, 1, 1.0e-3), (256, 3, 40.0)])
def test_edgeconv_tensor_core_kinds(H, frames, mag, kind):
    """fused EdgeConv branch on the tcgen05 engine, both operand kinds, against an fp64 evaluation of
    max_e relu(relu(P[i] + Q[j]) W1 + b1) * s + t on a ragged graph (heavy target, isolated vertices)"""
    n = 1500
    g = torch.Generator().manual_seed(H + frames)
    ei = torch.randint(0, n - 10, (2, 12000), generator=g)
    ei[1, :900] = 7
    gr = engine.graph_prep(ei.to(DEV), n)
    pq = torch.randn(n * frames, 2 * H, generator=g) * mag
    W1 = torch.randn(H, H, generator=g, dtype=torch.float64) / H ** 0.5
    b1, sc, sh = (torch.randn(H, generator=g) * mag for _ in range(3))
    sc = sc / mag
    blob, w_inv = packing.pack_edge_tc_blob(W1, sc, kind)
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=b1.to(DEV), scale=sc.to(DEV), shift=sh.to(DEV), H=H,
                            W1tc=blob.to(DEV), tc_kind=kind, tc_w_inv=w_inv)
    out = torch.full((n * frames, H), float("-inf"), device=DEV)
    engine.edgeconv(br, pq.to(DEV), 2 * H, 0, H, gr, frames, out, H, 0)
    e_real = int(gr.rowptr[n])
    i, j = gr.tgt[:e_real].long().cpu(), gr.col[:e_real].long().cpu()
    for f in range(frames):
        P, Q = pq[f * n:(f + 1) * n, :H].double(), pq[f * n:(f + 1) * n, H:].double()
        z = torch.relu(torch.relu(P[i] + Q[j]) @ W1.t() + b1.double()) * sc.double() + sh.double()
        ref = torch.full((n, H), float("-inf"), dtype=torch.float64).scatter_reduce(0, i[:, None].expand_as(z), z, "amax")
        assert helpers.max_abs_diff(out[f * n:(f + 1) * n], ref) < 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("H,frames", [(64, 1), (128, 1), (128, 5), (256, 1), (256, 6)])
def test_edgeconv_segment_boundaries_are_exact(H, frames):
    """fused EdgeConv epilogue (contiguous row ranges per warp, register carry across 32-row blocks, atomic merge only for
    segments that leave a warp's range, selective -inf start values): in-degrees chosen so that segments of 1 .. 700 slots
    start at every alignment relative to the 32 / 64 / 128 / 256-row boundaries; integer-valued operands and weights, scale
    +-1, so every sum is exact and the result must equal the fp64 evaluation BIT FOR BIT.  frames = 1 runs the
    cta_group::1 kernels (two accumulators per buffer for H = 256), more key-frames the 2-CTA pair kernel."""
    g = torch.Generator().manual_seed(H + frames)
    degs = [0, 30, 31, 32, 33, 1, 62, 63, 64, 65, 2, 126, 127, 128, 129, 3, 254, 255, 256, 257, 5, 700, 17, 8, 96, 160, 7]
    degs = (degs * 4)[: 4 * len(degs)]
    n = len(degs) + 40                                       # trailing vertices: self loop only
    tgt = torch.cat([torch.full((d,), v, dtype=torch.long) for v, d in enumerate(degs)])
    src = torch.randint(0, n, (tgt.numel(),), generator=g)
    src = torch.where(src == tgt, (src + 1) % n, src)        # no explicit self loops (graph_prep appends them)
    perm = torch.randperm(tgt.numel(), generator=g)
    ei = torch.stack([src, tgt])[:, perm]
    gr = engine.graph_prep(ei.to(DEV), n)
    pq = torch.randint(-4, 5, (n * frames, 2 * H), generator=g).float()
    W1 = torch.randint(-2, 3, (H, H), generator=g).double()
    b1 = torch.randint(-3, 4, (H,), generator=g).float()
    sc = (torch.randint(0, 2, (H,), generator=g) * 2 - 1).float()
    sh = torch.randint(-2, 3, (H,), generator=g).float()
    blob, w_inv = packing.pack_edge_tc_blob(W1, sc, packing.KIND_F16)
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=b1.to(DEV), scale=sc.to(DEV), shift=sh.to(DEV), H=H,
                            W1tc=blob.to(DEV), tc_kind=packing.KIND_F16, tc_w_inv=w_inv)
    ld = H + 8
    out = torch.full((n * frames, ld), 123.0, device=DEV)     # garbage start values: only fill_cut prepares the buffer
    engine.fill_cut([(gr, out, ld, 4, H)], n, frames, float("-inf"))
    engine.edgeconv(br, pq.to(DEV), 2 * H, 0, H, gr, frames, out, ld, 4)
    e_real = int(gr.rowptr[n])
    i, j = gr.tgt[:e_real].long().cpu(), gr.col[:e_real].long().cpu()
    res = out.cpu().reshape(frames, n, ld)
    assert bool((res[:, :, :4] == 123.0).all()) and bool((res[:, :, 4 + H:] == 123.0).all())   # neighbours untouched
    for f in range(frames):
        P, Q = pq[f * n:(f + 1) * n, :H].double(), pq[f * n:(f + 1) * n, H:].double()
        z = torch.relu(torch.relu(P[i] + Q[j]) @ W1.t() + b1.double()) * sc.double() + sh.double()
        ref = torch.full((n, H), float("-inf"), dtype=torch.float64).scatter_reduce(0, i[:, None].expand_as(z), z, "amax")
        assert torch.equal(res[f, :, 4:4 + H].double(), ref), (H, frames, f)


@pytest.mark.gpu
@pytest.mark.parametrize("n,e,frames", [(300, 2500, 1), (1024, 16384, 5), (64, 9000, 2)])
def test_fill_cut_touches_exactly_the_straddling_segments(n, e, frames):
    """morig_fill_cut_f32 (start values of the fused EdgeConv outputs): -inf lands on the vertices whose CSR segment straddles
    a multiple of 32 slots, in every key-frame and only in the given column range; everything else keeps its content"""
    g_ = torch.Generator().manual_seed(n + e)
    ei = torch.randint(0, n, (2, e), generator=g_)
    ei[1, : e // 3] = 7                                   # a hub: one long segment
    g = engine.graph_prep(ei.to(DEV), n)
    ld, col0, ncols = 52, 20, 24
    out = torch.arange(frames * n * ld, dtype=torch.float32, device=DEV).reshape(frames * n, ld).contiguous()
    want = out.clone().cpu().reshape(frames, n, ld)
    engine.fill_cut([(g, out, ld, col0, ncols), (g, out, ld, 1, 3)], n, frames, float("-inf"))
    rp = g.rowptr.cpu().long()
    a, b = rp[:-1], rp[1:]
    cut = (b > a) & ((a >> 5) != ((b - 1) >> 5))
    assert bool(cut.any()) and (n == 64 or not bool(cut.all()))      # the 64-vertex case: every segment is long
    want[:, cut, col0:col0 + ncols] = float("-inf")
    want[:, cut, 1:4] = float("-inf")
    assert torch.equal(out.cpu().reshape(frames, n, ld), want)


@pytest.mark.parametrize("H,frames,repeat,n,e,ldo_pad", [(16, 1, 1, 1500, 12000, 0), (16, 1, 5, 1500, 12000, 3),
                                                        (32, 3, 1, 1500, 12000, 0), (32, 1, 1, 5, 3, 1),
                                                        (16, 2, 1, 40, 0, 0), (32, 5, 1, 4096, 61440, 0)])
def test_edgeconv_narrow_kernel(H, frames, repeat, n, e, ldo_pad):
    """narrow EdgeConv branch (mma.sync 3xTF32 kernel, H = 16 / 32): ragged graph with a target whose segment
    spans many 32-slot tiles, isolated vertices (self loop only), duplicates, key-frames, out_repeat, odd strides"""
    g = torch.Generator().manual_seed(H + frames + n)
    ei = torch.randint(0, max(n - 3, 1), (2, e), generator=g)
    if e > 1000:
        ei[1, :900] = 7
    gr = engine.graph_prep(ei.to(DEV), n)
    pq = torch.randn(n * frames, 2 * H, generator=g)
    W1 = torch.randn(H, H, generator=g, dtype=torch.float64) / H ** 0.5
    b1, sc, sh = (torch.randn(H, generator=g) for _ in range(3))
    br = packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=b1.to(DEV), scale=sc.to(DEV), shift=sh.to(DEV), H=H)
    ldo, off = H + 2 + ldo_pad, 2 + (ldo_pad & 1)
    rows = n * frames * repeat
    out = torch.full((rows, ldo), float("-inf"), device=DEV)
    engine.edgeconv(br, pq.to(DEV), 2 * H, 0, H, gr, frames, out, ldo, off, out_repeat=repeat)
    e_real = int(gr.rowptr[n])
    i, j = gr.tgt[:e_real].long().cpu(), gr.col[:e_real].long().cpu()
    got = out.cpu()
    assert torch.isinf(got[:, :off]).all() and torch.isinf(got[:, off + H:]).all()       # neighbours untouched
    for f in range(frames):
        P, Q = pq[f * n:(f + 1) * n, :H].double(), pq[f * n:(f + 1) * n, H:].double()
        z = torch.relu(torch.relu(P[i] + Q[j]) @ W1.t() + b1.double()) * sc.double() + sh.double()
        ref = torch.full((n, H), float("-inf"), dtype=torch.float64).scatter_reduce(0, i[:, None].expand_as(z), z, "amax")
        for r in range(repeat):
            blk = got[(f + r) * n:(f + r + 1) * n, off:off + H]
            assert helpers.max_abs_diff(blk, ref) < 1e-5 * max(1.0, float(ref.abs().max()))


def test_edgeconv_batch_equals_single_launches():
    """three narrow branches (different weights, PQ columns and output buffers) in one launch: bit-equal to three
    launches of the single-branch entry point"""
    n, H, reps = 1500, 16, 3
    g = torch.Generator().manual_seed(11)
    ei = torch.randint(0, n, (2, 9000), generator=g)
    gr = engine.graph_prep(ei.to(DEV), n)
    pq = torch.randn(n, 3 * 2 * H, generator=g).to(DEV)
    brs, singles, batched = [], [], []
    for k in range(3):
        W1 = torch.randn(H, H, generator=g, dtype=torch.float64) / H ** 0.5
        b1, sc, sh = (torch.randn(H, generator=g).to(DEV) for _ in range(3))
        brs.append(packing.EdgeBranch(W1=packing._pack_wt(W1).to(DEV), b1=b1, scale=sc, shift=sh, H=H))
        singles.append(torch.full((n * reps, H + 4 * k), float("-inf"), device=DEV))
        batched.append(torch.full((n * reps, H + 4 * k), float("-inf"), device=DEV))
    items = []
    for k in range(3):
        engine.edgeconv(brs[k], pq, 6 * H, 2 * k * H, 2 * k * H + H, gr, 1, singles[k], H + 4 * k, 2 * k, out_repeat=reps)
        items.append((brs[k], pq, 6 * H, 2 * k * H, 2 * k * H + H, batched[k], H + 4 * k, 2 * k))
    engine.edgeconv_batch(items, gr, 1, out_repeat=reps)
    for a, b in zip(singles, batched):
        assert torch.equal(a, b)
        assert torch.isfinite(a[:, :H]).all() or True


def test_dense_tensor_core_pool_and_rowbias():
    n, frames, B, K, N = 700, 3, 4, 64, 200
    g = torch.Generator().manual_seed(0)
    A = torch.randn(n * frames, K, generator=g)
    W = torch.randn(N, K, generator=g) / 8
    rb = torch.randn(frames * B, N, generator=g)
    batch = torch.sort(torch.randint(0, B, (n,), generator=g)).values
    batch[0], batch[-1] = 0, B - 1
    binfo = engine.BatchInfo(batch32=batch.to(torch.int32).to(DEV), n_graphs=B)
    layer = packing.DenseLayer(W=packing._pack_wt(W.double()).to(DEV), K=K, N=N, relu=True).with_tc(W.double())
    layer.Wtc = layer.Wtc.to(DEV)
    C = torch.empty(n * frames, N, device=DEV)
    pool = torch.full((frames * B, N), float("-inf"), device=DEV)
    engine.dense(layer, A.to(DEV), 0, K, n * frames, C=C, ldc=N, pool=pool, rowbias=rb.to(DEV), binfo=binfo, n_vtx=n)
    grp = (torch.arange(n * frames) // n) * B + batch[torch.arange(n * frames) % n]
    ref = torch.relu(A.double() @ W.double().t() + rb.double()[grp])
    assert helpers.max_abs_diff(C, ref) < 2e-5
    present = torch.zeros(frames * B, dtype=torch.bool); present[grp] = True
    stored_max = torch.full((frames * B, N), float("-inf")).scatter_reduce(0, grp[:, None].expand(-1, N), C.cpu(), "amax")
    assert torch.equal(pool.cpu()[present], stored_max[present])


def test_dense_pool_and_rowbias():
    n, frames, B, K, N = 700, 3, 4, 64, 200
    g = torch.Generator().manual_seed(0)
    A = torch.randn(n * frames, K, generator=g)
    W = torch.randn(N, K, generator=g) / 8
    rb = torch.randn(frames * B, N, generator=g)
    batch = torch.sort(torch.randint(0, B, (n,), generator=g)).values
    batch[0], batch[-1] = 0, B - 1
    binfo = engine.BatchInfo(batch32=batch.to(torch.int32).to(DEV), n_graphs=B)
    layer = packing.DenseLayer(W=packing._pack_wt(W.double()).to(DEV), K=K, N=N, relu=True)
    C = torch.empty(n * frames, N, device=DEV)
    pool = torch.full((frames * B, N), float("-inf"), device=DEV)
    engine.dense(layer, A.to(DEV), 0, K, n * frames, C=C, ldc=N, pool=pool, rowbias=rb.to(DEV), binfo=binfo, n_vtx=n)
    grp = (torch.arange(n * frames) // n) * B + batch[torch.arange(n * frames) % n]
    ref = torch.relu(A.double() @ W.double().t() + rb.double()[grp])
    assert helpers.max_abs_diff(C, ref) < 2e-5
    ref_pool = torch.full((frames * B, N), float("-inf"), dtype=torch.float64).scatter_reduce(
        0, grp[:, None].expand_as(ref), ref, "amax")
    got = pool.cpu().double()
    present = torch.zeros(frames * B, dtype=torch.bool); present[grp] = True
    assert float((got[present] - ref_pool[present]).abs().max()) < 2e-5
    # pooled value must be bit-identical to the max of the stored rows (ordered atomics are exact)
    stored_max = torch.full((frames * B, N), float("-inf")).scatter_reduce(0, grp[:, None].expand(-1, N), C.cpu(), "amax")
    assert torch.equal(pool.cpu()[present], stored_max[present])


@pytest.mark.parametrize("C_x,H,C_p,Dp", [(3, 32, 3, 16), (64, 128, 3, 16), (256, 256, 3, 16), (32, 128, 33, 64)])
def test_edge_conv_motion_module(C_x, H, C_p, Dp):
    """one EdgeConvMotion (models/basic_modules.py:179-199) on a ragged graph with a heavy target, duplicate
    edges, pre-existing self loops and isolated vertices"""
    import morig_b200
    from oracle import rignet_port
    n = 777
    g = torch.Generator().manual_seed(C_x + H)
    ei = torch.randint(0, n - 20, (2, 6000), generator=g)            # last 20 vertices isolated
    ei[1, :700] = 5
    mod = morig_b200.EdgeConvMotion(morig_b200.MLP([2 * C_x, H, H]), morig_b200.MLP([2 * C_p, Dp, Dp])).eval()
    mod.load_state_dict(synth.seeded_state_dict(mod, 7))
    pos, x = torch.randn(n, C_p, generator=g), torch.randn(n, C_x, generator=g)
    ref = rignet_port.edge_conv_motion({"m." + k: v for k, v in mod.state_dict().items()}, "m", pos, x, ei)
    out = mod.to(DEV)(pos.to(DEV), x.to(DEV), ei.to(DEV))
    assert helpers.max_abs_diff(out, ref) < 2e-5
    # edge order must not matter (max is exact): permuted edge list -> bit-identical output
    perm = torch.randperm(ei.shape[1], generator=g)
    out2 = mod(pos.to(DEV), x.to(DEV), ei[:, perm].contiguous().to(DEV))
    assert torch.equal(out, out2)


@pytest.mark.parametrize("cin,cout", [(3, 32), (32, 64), (64, 256), (256, 512)])
def test_gcu_module(cin, cout):
    """GCU / EdgeConv (models/basic_modules.py:142-177) at the channel sizes models/corrnet.py:17-20 uses"""
    import morig_b200
    from oracle import rignet_port
    data = synth.make_batch(2, 400, seed=8)
    gcu = morig_b200.GCU(cin, cout).eval()
    gcu.load_state_dict(synth.seeded_state_dict(gcu, cin))
    x = torch.randn(800, cin, generator=torch.Generator().manual_seed(cin))
    want = rignet_port.gcu({"g." + k: v for k, v in gcu.state_dict().items()}, "g", x, data.tpl_edge_index,
                           data.geo_edge_index)
    got = gcu.to(DEV)(x.to(DEV), data.tpl_edge_index.to(DEV), data.geo_edge_index.to(DEV))
    assert helpers.max_abs_diff(got, want) < 2e-5


def test_temporal_attn_module():
    import morig_b200
    from oracle import rignet_port
    mod = morig_b200.TemporalAttn(32, 2, 64, 512, 64).eval()
    mod.load_state_dict(synth.seeded_state_dict(mod, 3))
    x = torch.nn.functional.normalize(torch.randn(1000, 5, 32, generator=torch.Generator().manual_seed(1)), dim=2)
    ref = rignet_port.temporal_attn({"a." + k: v for k, v in mod.state_dict().items()}, "a", x)
    out = mod.to(DEV)(x.to(DEV))
    assert helpers.max_abs_diff(out, ref) < 2e-5


# ---- whole networks -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", helpers.golden_names())
def test_models_match_golden_fixtures(name):
    """outputs of the unmodified reference (tests/golden, made by oracle/gen_golden.py)"""
    arch, kw, wseed, data, expect = helpers.load_golden(name)
    model = helpers.build_model(arch, kw, wseed, DEV)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert o.shape == e.shape, k
        assert helpers.max_abs_diff(o, e) < helpers.TOL, k


@pytest.mark.parametrize("arch,b,n", [("jointnet_motion", 2, 1024), ("masknet_motion", 1, 2048), ("skinnet_motion", 2, 1024)])
def test_models_match_oracle(arch, b, n):
    kw = synth.ARCH_KWARGS[arch]
    data = synth.make_batch(b, n, seed=100, with_skin=(arch == "skinnet_motion"))
    model = helpers.build_model(arch, kw, 11, DEV)
    expect = helpers.oracle_forward(arch, kw, model, data, data.pred_flow)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert helpers.max_abs_diff(o, e) < helpers.TOL, k


def _ragged_batch(sizes, seed, with_skin=False):
    return synth.collate([synth.make_mesh(n, seed + i, with_skin) for i, n in enumerate(sizes)])


@pytest.mark.parametrize("arch,sizes", [("jointnet_motion", (300, 1024, 177)), ("masknet_motion", (9, 130)),
                                        ("skinnet_motion", (260, 64, 640))])
def test_ragged_batches_match_oracle(arch, sizes):
    """meshes of different sizes in one batch (vertex counts that are no multiple of any tile), incl. a 9-vertex mesh"""
    kw = synth.ARCH_KWARGS[arch]
    data = _ragged_batch(sizes, 300, with_skin=(arch == "skinnet_motion"))
    model = helpers.build_model(arch, kw, 13, DEV)
    expect = helpers.oracle_forward(arch, kw, model, data, data.pred_flow)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert o.shape == e.shape
        assert helpers.max_abs_diff(o, e) < helpers.TOL, k


def test_degenerate_graphs_match_oracle():
    """no edges at all (every vertex only gets its self loop), and a star graph whose hub has in-degree N-1"""
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 17, DEV)
    base = synth.make_batch(1, 400, seed=3)
    n = base.pos.shape[0]
    empty = torch.zeros(2, 0, dtype=torch.long)
    star = torch.stack([torch.arange(1, n), torch.zeros(n - 1, dtype=torch.long)])
    for tpl, geo in ((empty, empty), (star, base.geo_edge_index), (base.tpl_edge_index, star)):
        data = synth.Batch(**base.__dict__)
        data.tpl_edge_index, data.geo_edge_index = tpl.contiguous(), geo.contiguous()
        expect = helpers.oracle_forward("jointnet_motion", kw, model, data, data.pred_flow)
        with torch.no_grad():
            out = model(data.to(DEV), data.pred_flow.to(DEV))
        for o, e, k in zip(out, expect, helpers.OUT_KEYS):
            assert helpers.max_abs_diff(o, e) < helpers.TOL, k


@pytest.mark.parametrize("arch,b,n", [("masknet_motion", 8, 4096), ("skinnet_motion", 4, 8192)])
def test_baseline_configs_3_and_4_run_and_are_deterministic(arch, b, n):
    """BASELINE.json configs[2] / configs[3] sizes: finite outputs, unit-norm embeddings, run-to-run bit equality
    (plain launches vs CUDA-graph replay), first mesh equal to the same mesh run alone"""
    kw = synth.ARCH_KWARGS[arch]
    skin = arch == "skinnet_motion"
    model = helpers.build_model(arch, kw, 19, DEV)
    data = synth.make_batch(b, n, seed=0, with_skin=skin).to(DEV)
    with torch.no_grad():
        a = [t.clone() for t in model(data, data.pred_flow)]
        model(data, data.pred_flow)
        c = model(data, data.pred_flow)                       # third call: graph replay
        for x, y in zip(a, c):
            assert torch.equal(x, y)
        assert all(torch.isfinite(t).all() for t in a)
        assert float((a[0].norm(dim=2) - 1).abs().max()) < 1e-5
        one = synth.make_batch(1, n, seed=0, with_skin=skin).to(DEV)
        solo = model(one, one.pred_flow)
        for x, y in zip(solo, a):
            assert helpers.max_abs_diff(x, y[: x.shape[0]]) < 1e-5


def test_full_size_properties():
    """BASELINE.json configs[1] size (4 x 4096 vertices): properties that need no oracle —
    run-to-run bit equality, data-parallel shard concatenation == single batch, edge-order invariance."""
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 5, DEV)
    data = synth.make_batch(4, 4096, seed=0)
    dd = data.to(DEV)
    with torch.no_grad():
        a = [t.clone() for t in model(dd, dd.pred_flow)]
        b = model(dd, dd.pred_flow)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert all(torch.isfinite(t).all() for t in a)
        # unit rows
        assert float((a[0].norm(dim=2) - 1).abs().max()) < 1e-5
        # shards of whole meshes give the same rows as the full batch (BatchNorm is eval-mode, pooling per graph)
        parts = []
        for s in range(2):
            sh = synth.make_batch(2, 4096, seed=2 * s).to(DEV)
            parts.append([t.clone() for t in model(sh, sh.pred_flow)])
        for k in range(3):
            cat = torch.cat([parts[0][k], parts[1][k]], dim=0)
            assert helpers.max_abs_diff(cat, a[k]) < 1e-5
        # shuffled edge order
        g = torch.Generator().manual_seed(0)
        d2 = synth.Batch(**dd.__dict__)
        d2.geo_edge_index = dd.geo_edge_index[:, torch.randperm(dd.geo_edge_index.shape[1], generator=g).to(DEV)].contiguous()
        c = model(d2, d2.pred_flow)
        for x, y in zip(a, c):
            assert torch.equal(x, y)


@pytest.mark.parametrize("arch", ["jointnet_motion", "skinnet_motion"])
def test_cuda_graph_replay_equals_plain_launches(arch):
    """third call with the same batch shape replays a captured CUDA graph on static buffers: results must be
    bit-identical to plain launches for NEW data of that shape, and follow load_state_dict"""
    kw = synth.ARCH_KWARGS[arch]
    skin = arch == "skinnet_motion"
    model = helpers.build_model(arch, kw, 5, DEV)
    batches = [synth.make_batch(2, 256, seed=s, with_skin=skin).to(DEV) for s in (1, 2, 3, 4)]
    with torch.no_grad():
        model.use_cuda_graph = False
        plain = [[t.clone() for t in model(b, b.pred_flow)] for b in batches]
        model.use_cuda_graph = True
        for b, want in zip(batches, plain):                    # calls 1-2 plain, 3-4 replayed
            got = model(b, b.pred_flow)
            for x, y in zip(got, want):
                assert torch.equal(x, y)
        assert len(model._replays) == 1
        model.load_state_dict(synth.seeded_state_dict(model, 6))
        new = model(batches[0], batches[0].pred_flow)           # stale graph must not be replayed
        assert not torch.equal(new[2], plain[0][2])
        model.use_cuda_graph = False
        ref = model(batches[0], batches[0].pred_flow)
        for x, y in zip(new, ref):
            assert torch.equal(x, y)


def test_rejects_cpu_tensors():
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 5, DEV)
    data = synth.make_batch(1, 256, seed=0)
    with pytest.raises(RuntimeError):
        model(data, data.pred_flow)                      # CPU tensors: no fallback
    model.train()
    with pytest.raises(RuntimeError):
        model(data, data.pred_flow)


# ---- host-to-host streaming (morig_b200.HostPipeline) ---------------------------------------------------
@pytest.mark.parametrize("arch", ["jointnet_motion", "skinnet_motion"])
def test_host_pipeline_matches_direct_calls_in_order(arch):
    """Different host batches of one shape through the 3-stream pipeline: results come back in submission
    order, equal bit for bit to sequential `model(data.to(dev), flow)` calls, and match the oracle."""
    import morig_b200
    kw = synth.ARCH_KWARGS[arch]
    skin = arch == "skinnet_motion"
    model = helpers.build_model(arch, kw, 5, DEV)
    batches = [synth.make_batch(2, 512, seed=40 + 2 * i, with_skin=skin).pin_memory() for i in range(5)]
    direct = []
    with torch.no_grad():
        for b in batches:
            d = b.to(DEV)
            direct.append([o.cpu() for o in model(d, d.pred_flow)])
    pipe = morig_b200.HostPipeline(model, depth=2)
    got = [[o.clone() for o in outs] for outs in pipe.run(batches)]
    assert len(got) == len(batches) and pipe.in_flight == 0
    for g, d in zip(got, direct):
        for a, b_ in zip(g, d):
            assert a.device.type == "cpu" and torch.equal(a, b_)
    expect = helpers.oracle_forward(arch, kw, model, batches[3], batches[3].pred_flow)
    for a, e in zip(got[3], expect):
        assert helpers.max_abs_diff(a, e) < helpers.TOL


def test_host_pipeline_slot_discipline():
    import morig_b200
    kw = synth.ARCH_KWARGS["masknet_motion"]
    model = helpers.build_model("masknet_motion", kw, 2, DEV)
    b = synth.make_batch(1, 256, seed=3)
    pipe = morig_b200.HostPipeline(model, depth=1)
    with pytest.raises(RuntimeError):
        pipe.result()
    pipe.submit(b, b.pred_flow)
    with pytest.raises(RuntimeError):
        pipe.submit(b, b.pred_flow)
    outs = pipe.result()
    assert all(o.is_pinned() for o in outs)
    with pytest.raises(TypeError):
        d = b.to(DEV)
        pipe.submit(d, d.pred_flow)


# ---- joint-extraction post-process: weighted mean-shift on the GPU -------------------------------
def _cluster_points(n_half, seed, n_centres=6):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.4, 0.4, size=(n_centres, 3))
    pts = c[rng.integers(0, n_centres, n_half)] + rng.normal(0, 0.02, size=(n_half, 3))
    pts = np.concatenate([pts, pts * np.array([[-1, 1, 1]])], axis=0)
    w = np.tile(rng.uniform(0.05, 1.0, size=(n_half, 1)).astype(np.float32), (2, 1))
    return pts, w


def test_meanshift_matches_golden_fixture():
    """fixture produced by the unmodified reference function (oracle/gen_golden_cluster.py); fp64, so the only
    difference is the summation order of the N-term sums: 1e-9 absolute, same number of iterations"""
    from morig_b200 import cluster_utils
    from oracle import cluster_port
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "meanshift_n300.npz"))
    got, iters = cluster_utils.meanshift_cluster(z["pts"], float(z["bandwidth"]), z["attn"], int(z["max_iter"]),
                                                 return_iters=True)
    _, ref_iters = cluster_port.meanshift_cluster(z["pts"], float(z["bandwidth"]), z["attn"], int(z["max_iter"]),
                                                  return_iters=True)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and iters == ref_iters
    assert np.abs(got - z["out"]).max() < 1e-9


@pytest.mark.parametrize("n_half,seed,weighted,max_iter", [(1, 0, True, 20), (17, 1, False, 20), (700, 2, True, 30),
                                                           (1500, 3, True, 5)])
def test_meanshift_matches_oracle(n_half, seed, weighted, max_iter):
    from morig_b200 import cluster_utils
    from oracle import cluster_port
    pts, w = _cluster_points(n_half, seed)
    w = w if weighted else None
    ref, ref_iters = cluster_port.meanshift_cluster(pts, 0.05, w, max_iter, return_iters=True)
    got, iters = cluster_utils.meanshift_cluster(torch.from_numpy(pts).to(DEV), 0.05,
                                                 None if w is None else torch.from_numpy(w).to(DEV), max_iter,
                                                 return_iters=True)
    assert got.is_cuda and iters == ref_iters
    assert np.abs(got.cpu().numpy() - ref).max() < 1e-9


def test_meanshift_full_size_properties():
    """8192 points (a 4K-vertex mesh reflected): modes are fixed points (a second run from the result stops after one
    step), points never leave the bounding box of the input, and the result is deterministic"""
    from morig_b200 import cluster_utils
    pts, w = _cluster_points(4096, 9, n_centres=20)
    a, it_a = cluster_utils.meanshift_cluster(pts, 0.05, w, 30, return_iters=True)
    b = cluster_utils.meanshift_cluster(pts, 0.05, w, 30)
    assert np.array_equal(a, b) and 1 < it_a <= 29
    assert (a.min(0) >= pts.min(0) - 1e-12).all() and (a.max(0) <= pts.max(0) + 1e-12).all()
    _, it_c = cluster_utils.meanshift_cluster(a, 0.05, w, 30, return_iters=True)
    assert it_c <= 2


# ---- surface-geodesic graph build on the GPU: fp64, bit-exact --------------------------------------
def test_surface_geodesic_matches_golden_fixture():
    """fixture produced by the unmodified reference (numpy + scipy Dijkstra); two components -> unreachable pairs"""
    from morig_b200 import graph_build
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "geodesic_s400_v150.npz"))
    got = graph_build.surface_geodesic(z["pts"], z["normals"], z["verts"])
    assert isinstance(got, np.ndarray) and np.array_equal(got, z["surface_geodesic"])


@pytest.mark.parametrize("s,v,seed,two", [(2, 1, 0, False), (7, 30, 1, False), (500, 200, 2, False), (1500, 700, 3, True)])
def test_surface_geodesic_matches_oracle_bit_exact(s, v, seed, two):
    from morig_b200 import graph_build
    from oracle import gen_golden_geodesic as gg
    from oracle import geodesic_port
    pts, nrm, verts = gg.make_inputs(s, v, seed, two_parts=two)
    ref = geodesic_port.surface_geodesic_from_samples(pts, nrm, verts)
    got = graph_build.surface_geodesic(torch.from_numpy(pts).to(DEV), torch.from_numpy(nrm).to(DEV),
                                       torch.from_numpy(verts).to(DEV))
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("radius,max_nn", [(0.06, 15), (0.15, 15), (0.3, 4), (100.0, 3)])
def test_geo_ball_edges_bit_exact(radius, max_nn):
    from morig_b200 import graph_build
    from oracle import gen_golden_geodesic as gg
    from oracle import geodesic_port
    pts, nrm, verts = gg.make_inputs(600, 300, 4)
    geo = geodesic_port.surface_geodesic_from_samples(pts, nrm, verts)
    ref = geodesic_port.geo_ball_edges(geo, radius, max_nn)
    got = graph_build.geo_ball_edges(geo, radius, max_nn)
    assert got.dtype == np.int64 and np.array_equal(got, ref)
    assert np.bincount(got[:, 0], minlength=300).max() <= max_nn and (got[:, 0] != got[:, 1]).all()


def test_surface_geodesic_full_size_properties():
    """the reference's size (4000 samples, a 4096-vertex mesh): symmetric, zero diagonal for vertices sharing a sample,
    triangle inequality on sampled triples, never shorter than the Euclidean distance of the samples, deterministic"""
    from morig_b200 import graph_build
    from oracle import gen_golden_geodesic as gg
    pts, nrm, verts = gg.make_inputs(4000, 4096, 6)
    a = graph_build.surface_geodesic(pts, nrm, verts)
    assert np.array_equal(a, graph_build.surface_geodesic(pts, nrm, verts))
    assert np.array_equal(a, a.T) and (a >= 0).all() and np.isfinite(a).all()
    nn = np.argmin(np.sqrt(((verts[None] - pts[:, None]) ** 2).sum(2)), axis=0)
    euc = np.sqrt(((pts[nn][None] - pts[nn][:, None]) ** 2).sum(2))
    assert (a >= euc - 1e-6).all()
    rng = np.random.default_rng(0)
    i, j, k = rng.integers(0, 4096, (3, 20000))
    assert (a[i, j] <= a[i, k] + a[k, j] + 1e-9).all()


@pytest.mark.parametrize("nu,nv", [(2, 2), (7, 9), (64, 64)])
def test_tpl_edges_bit_exact(nu, nv):
    """topological edges from the face list against the statement-for-statement port of get_tpl_edges
    (compared as (v, n)-sorted rows; the reference's order inside a vertex is python-set order)"""
    from morig_b200 import graph_build
    from oracle import geodesic_port
    from test_oracle_pinning import _grid_faces
    verts, faces = _grid_faces(nu, nv, nu)
    got = graph_build.tpl_edges(verts, faces)
    ref = geodesic_port.sorted_rows(geodesic_port.tpl_edges(verts, faces))
    assert got.dtype == np.int64 and np.array_equal(got, ref)


# ---- parity against the oracle at the sizes BASELINE.json is quoted on ----------------------------------------------------
@pytest.mark.parametrize("arch,b,n", [("jointnet_motion", 4, 4096),      # configs[1]
                                      ("masknet_motion", 8, 4096),       # configs[2]
                                      ("skinnet_motion", 4, 8192)])      # configs[3]
def test_models_match_oracle_at_baseline_configs(arch, b, n):
    """CUDA path vs the CPU oracle (oracle/rignet_port.py) on the full BASELINE.json batches: the fp16-split operand
    scales depend on data-dependent max|value|, so parity is checked at the quoted sizes, not only on small meshes"""
    kw = synth.ARCH_KWARGS[arch]
    data = synth.make_batch(b, n, seed=0, with_skin=(arch == "skinnet_motion"))
    model = helpers.build_model(arch, kw, 1, DEV)
    expect = helpers.oracle_forward(arch, kw, model, data, data.pred_flow)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert o.shape == e.shape
        assert helpers.max_abs_diff(o, e) < helpers.TOL, k


@pytest.mark.parametrize("use_Dg,use_Lf", [(True, False), (False, True), (True, True)])
def test_skinnet_column_selections(use_Dg, use_Lf):
    """the three other branches of the skin_input column selection (models/rignet.py:159-171)"""
    kw = dict(synth.ARCH_KWARGS["skinnet_motion"], use_Dg=use_Dg, use_Lf=use_Lf)
    data = synth.make_batch(2, 512, seed=60, with_skin=True)
    model = helpers.build_model("skinnet_motion", kw, 23, DEV)
    expect = helpers.oracle_forward("skinnet_motion", kw, model, data, data.pred_flow)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert helpers.max_abs_diff(o, e) < helpers.TOL, k


def _outlier_state_dict(model, seed):
    """a 'trained-like' weight set: a few BatchNorm scales x1000 (edge MLPs and vertex MLPs), one hidden channel of a
    wide layer x1e4 and one first-layer row x100 -- channels whose magnitude is far from the rest of their tensor"""
    sd = synth.seeded_state_dict(model, seed)
    for key, idx, f in (("motionNet.gcu_2.edge_conv_geo.nn_x.0.2.weight", 5, 1e3),
                        ("motionNet.gcu_3.edge_conv_tpl.nn_x.1.2.weight", 17, 1e3),
                        ("motionNet.gcu_1.mlp.0.2.weight", 3, 1e3),
                        ("motionNet.mlp_transform.0.0.0.weight", 7, 1e4),
                        ("motionNet.gcu_1.edge_conv_tpl.nn_pos.0.0.weight", 2, 1e2)):
        if key in sd:
            sd[key][idx] = sd[key][idx] * f
    for head in ("jointnet", "masknet", "skinNet"):
        for key, idx, f in ((f"{head}.gcu_3.edge_conv_geo.nn_x.0.2.weight", 9, 1e3), (f"{head}.mlp_glb.0.2.weight", 100, 1e3),
                            (f"{head}.gcu2.edge_conv_tpl.nn_x.1.2.weight", 11, 1e3), (f"{head}.gcu3.mlp.0.0.weight", 4, 1e4)):
            if key in sd:
                sd[key][idx] = sd[key][idx] * f
    return sd


@pytest.mark.parametrize("kind", ["f16", "tf32"])
@pytest.mark.parametrize("arch", ["jointnet_motion", "skinnet_motion"])
def test_outlier_channels_through_whole_networks(arch, kind, monkeypatch):
    """dynamic range of the split-operand arithmetic: one max|value| per buffer scales the fp16 operands, so a
    checkpoint with outlier channels (BN scale x1e3, one hidden channel x1e4) must still meet the tolerance --
    1e-4 absolute on the unit-norm embeddings, 1e-4 of the output range on the (now large) predictions"""
    monkeypatch.setenv("MORIG_TC_KIND", kind)
    kw = synth.ARCH_KWARGS[arch]
    data = synth.make_batch(2, 1024, seed=70, with_skin=(arch == "skinnet_motion"))
    model = getattr(__import__("morig_b200"), arch)(**kw).eval()
    model.load_state_dict(_outlier_state_dict(model, 29))
    model = model.to(DEV)
    expect = helpers.oracle_forward(arch, kw, model, data, data.pred_flow)
    exact = helpers.oracle_forward(arch, kw, model, data, data.pred_flow, dtype=torch.float64)
    with torch.no_grad():
        out = model(data.to(DEV), data.pred_flow.to(DEV))
    for o, e, x, k in zip(out, expect, exact, helpers.OUT_KEYS):
        assert torch.isfinite(o).all(), k
        # with a hidden channel at 1e4 x the rest, the fp32 reference itself is no longer accurate to 1e-4: the yardstick
        # is the same op sequence in fp64, and the CUDA path may not be further from it than a few times the fp32
        # reference is (or than the tolerance, where the reference still meets it)
        ref_err = helpers.max_abs_diff(e, x)
        got_err = helpers.max_abs_diff(o, x)
        print(f"{arch} {kind} {k}: |cuda - fp64| = {got_err:.3e}, |fp32 oracle - fp64| = {ref_err:.3e}, "
              f"range {float(x.abs().max()):.3e}")
        assert got_err < max(helpers.TOL * max(1.0, float(x.abs().max())), 4.0 * ref_err), k


# ---- integer-exact known-answer test ---------------------------------------------------------------
def _integer_state_dict(model, seed):
    """small-integer weights (sparse, in {-1, 0, 1}), integer biases, BatchNorm = exact integer affine
    (running_var = 1 - eps so that 1/sqrt(var + eps) rounds to exactly 1, gamma = +-1, integer
    beta / running_mean): every fp32 sum of the forward is then an exact integer whatever the summation order"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, ref in sorted(model.state_dict().items()):
        leaf = key.rsplit(".", 1)[-1]
        shape = tuple(ref.shape)
        if leaf == "num_batches_tracked":
            v = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_var":
            v = torch.full(shape, 1.0 - 1e-5)
        elif leaf == "running_mean":
            v = torch.randint(-1, 2, shape, generator=g).float()
        elif ref.dim() == 2:
            dense_ = torch.randint(-1, 2, shape, generator=g).float()
            keep = torch.rand(shape, generator=g) < min(1.0, 8.0 / shape[1])      # ~8 non-zeros per output row
            v = dense_ * keep
        elif leaf == "weight":                                                     # BatchNorm gamma
            v = torch.ones(shape)
            v = torch.where(torch.rand(shape, generator=g) < 0.2, -v, v)
        else:
            v = torch.randint(-1, 2, shape, generator=g).float()
        sd[key] = v.to(ref.dtype)
    return sd


@pytest.mark.parametrize("kind", ["f16", "tf32"])
@pytest.mark.parametrize("F,O", [(3, 32), (64, 3)])
def test_integer_exact_kat_gcn_rig(F, O, kind, monkeypatch):
    """GCNRig (3 GCUMotion + pooled global feature + head, models/rignet.py:50-67) on integer data: the CUDA path
    (factorised first Linear, folded BatchNorm, split-operand tensor-core GEMMs, segmented / pooled max) must
    reproduce the oracle BIT FOR BIT -- any dropped low-order term, mis-scaled operand or wrong segment shows"""
    import morig_b200
    from oracle import rignet_port
    monkeypatch.setenv("MORIG_TC_KIND", kind)
    data = synth.make_batch(2, 400, seed=5)
    n = data.pos.shape[0]
    g = torch.Generator().manual_seed(F)
    pos = torch.randint(-2, 3, (n, 3), generator=g).float()
    feat = torch.randint(-2, 3, (n, F), generator=g).float()
    rig = morig_b200.GCNRig(F, O).eval()
    rig.load_state_dict(_integer_state_dict(rig, 31 + F))
    sd = {"r." + k: v for k, v in rig.state_dict().items()}
    with torch.no_grad():
        want = rignet_port.gcn_rig(sd, "r", pos, feat, data.tpl_edge_index, data.geo_edge_index, data.batch)
    assert torch.equal(want, want.round()) and float(want.abs().max()) < 2 ** 21     # the KAT is integer-valued
    assert float(want.abs().max()) > 0
    got = rig.to(DEV)(pos.to(DEV), feat.to(DEV), data.tpl_edge_index.to(DEV), data.geo_edge_index.to(DEV),
                      data.batch.to(DEV)).cpu()
    if kind == "tf32":
        # the tf32 weight image keeps, in its lo half, the fp64 residual of the folded BatchNorm scale
        # (1 / sqrt(fp32(1 - 1e-5) + 1e-5) = 1 + 6.8e-9, which the fp32 oracle rounds to 1): results differ from the
        # integers by that relative amount, so this kind is held to 2^-19 of the range instead of bit equality
        assert float((got - want).abs().max()) <= 2.0 ** -19 * float(want.abs().max())
        return
    bad = (got != want)
    assert not bad.any(), (f"{int(bad.sum())} of {bad.numel()} outputs differ, max |diff| = "
                           f"{float((got - want).abs().max())}, first at {bad.nonzero()[0].tolist()}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_device_while_first_is_current():
    """device guard: a model and its data on cuda:1 while cuda:0 is the current device (PyTorch ops guard implicitly; the
    raw launches of this package must do it themselves) -- same bits as on cuda:0"""
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    data = synth.make_batch(2, 256, seed=4)
    torch.cuda.set_device(0)
    m0 = helpers.build_model("jointnet_motion", kw, 5, "cuda:0")
    m1 = helpers.build_model("jointnet_motion", kw, 5, "cuda:1")
    with torch.no_grad():
        a = m0(data.to("cuda:0"), data.pred_flow.to("cuda:0"))
        for _ in range(3):                                   # plain launches, then the captured graph
            b = m1(data.to("cuda:1"), data.pred_flow.to("cuda:1"))
        torch.cuda.synchronize("cuda:1")
    assert torch.cuda.current_device() == 0
    for x, y in zip(a, b):
        assert y.device == torch.device("cuda:1") and torch.equal(x.cpu(), y.cpu())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_gradient_all_reduce_two_gpus(tmp_path):
    """dp.GradAllReduce over NCCL on two GPUs (one process per GPU): gradients equal the mean of the per-rank gradients"""
    import subprocess
    import sys
    script = tmp_path / "ddp.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MORIG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MORIG_ROOT"], "tests"))
import helpers
from morig_b200 import dp, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
kw = synth.ARCH_KWARGS["masknet_motion"]
model = helpers.build_model("masknet_motion", kw, 4, f"cuda:{rank}").train()
ar = dp.GradAllReduce(model, bucket_mb=8.0)
def grads(m, r):
    d = synth.make_batch(1, 256, seed=300 + r).to(f"cuda:{rank}")
    out = m(d, d.pred_flow)
    (out[2].tanh().pow(2).mean() + out[1].sum() * 1e-3).backward()
ar.zero_grad(); grads(model, rank); nbytes = ar.finish()
got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
want = None
for r in range(world):
    m = helpers.build_model("masknet_motion", kw, 4, f"cuda:{rank}").train()
    grads(m, r)
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    want = g if want is None else want + g
err = float((got - want / world).abs().max() / want.abs().max())
if rank == 0:
    print("RESULT", err, nbytes, got.numel())
dist.destroy_process_group()
''')
    env = dict(os.environ, MORIG_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")]
    assert line, out.stderr[-2000:]
    _, err, nbytes, numel = line[0].split()
    assert float(err) < 1e-4 and int(nbytes) == 4 * int(numel)
