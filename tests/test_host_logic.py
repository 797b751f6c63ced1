"""Host-side logic on CPU (no GPU in the build container): weight folding / packing, buffer layouts and
launch sequences are exercised through the CPU emulation of the C-ABI in tests/emu.py and compared with
the golden outputs of the unmodified reference; plus the C-ABI library's exports and the module protocol
(state_dict, cache invalidation, loud failures)."""
import ctypes
import os
import re

import pytest
import torch

import helpers
from morig_b200 import _lib, engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", helpers.golden_names())
def test_packing_and_launch_sequence_reproduce_golden(emulated, name):
    arch, kw, wseed, data, expect = helpers.load_golden(name)
    model = helpers.build_model(arch, kw, wseed)
    with torch.no_grad():
        out = model(data, data.pred_flow)
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert o.shape == e.shape
        assert helpers.max_abs_diff(o, e) < 5e-6, k


def test_submodules_keep_reference_signatures(emulated):
    """EdgeConvMotion / GCUMotion / GCNRig / TemporalAttn called the way the reference calls them
    (models/basic_modules.py:185,214; models/rignet.py:36,58)"""
    from oracle import rignet_port
    import morig_b200
    data = synth.make_batch(2, 100, seed=5)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(200, 64, generator=g)
    gcu = morig_b200.GCUMotion(64, 256).eval()
    gcu.load_state_dict(synth.seeded_state_dict(gcu, 1))
    sd = {"g." + k: v for k, v in gcu.state_dict().items()}
    want = rignet_port.gcu_motion(sd, "g", data.pos, x, data.tpl_edge_index, data.geo_edge_index)
    got = gcu(data.pos, x, data.tpl_edge_index, data.geo_edge_index)
    assert helpers.max_abs_diff(got, want) < 5e-6
    ec = gcu.edge_conv_geo
    want = rignet_port.edge_conv_motion(sd, "g.edge_conv_geo", data.pos, x, data.geo_edge_index)
    assert helpers.max_abs_diff(ec(data.pos, x, data.geo_edge_index), want) < 5e-6
    rig = morig_b200.GCNRig(64, 3).eval()
    rig.load_state_dict(synth.seeded_state_dict(rig, 2))
    sd = {"r." + k: v for k, v in rig.state_dict().items()}
    want = rignet_port.gcn_rig(sd, "r", data.pos, x, data.tpl_edge_index, data.geo_edge_index, data.batch)
    got = rig(data.pos, x, data.tpl_edge_index, data.geo_edge_index, data.batch)
    assert helpers.max_abs_diff(got, want) < 5e-6
    att = morig_b200.TemporalAttn(32, 2, 64, 512, 64).eval()
    att.load_state_dict(synth.seeded_state_dict(att, 3))
    xs = torch.randn(50, 5, 32, generator=g)
    want = rignet_port.temporal_attn({"a." + k: v for k, v in att.state_dict().items()}, "a", xs)
    assert helpers.max_abs_diff(att(xs), want) < 5e-6


def test_gcu_and_edge_conv_without_pos_branch(emulated):
    """EdgeConv / GCU (models/basic_modules.py:142-177): the C_p = 0 case of the fused kernel"""
    from oracle import rignet_port
    import morig_b200
    data = synth.make_batch(2, 100, seed=8)
    for cin, cout in ((3, 32), (32, 64), (64, 256)):
        gcu = morig_b200.GCU(cin, cout).eval()
        gcu.load_state_dict(synth.seeded_state_dict(gcu, cin))
        x = torch.randn(200, cin, generator=torch.Generator().manual_seed(cin))
        sd = {"g." + k: v for k, v in gcu.state_dict().items()}
        want = rignet_port.gcu(sd, "g", x, data.tpl_edge_index, data.geo_edge_index)
        assert helpers.max_abs_diff(gcu(x, data.tpl_edge_index, data.geo_edge_index), want) < 5e-6
        want = rignet_port.edge_conv(sd, "g.edge_conv_tpl", x, data.tpl_edge_index)
        assert helpers.max_abs_diff(gcu.edge_conv_tpl(x, data.tpl_edge_index), want) < 5e-6


def test_negative_bn_scale_is_not_commuted_through_max(emulated):
    """BN after ReLU with gamma < 0: max(s*h+t) != s*max(h)+t.  All second-layer BN scales negative."""
    from oracle import rignet_port
    import morig_b200
    data = synth.make_batch(1, 64, seed=2)
    ec = morig_b200.EdgeConvMotion(morig_b200.MLP([6, 32, 32]), morig_b200.MLP([6, 16, 16])).eval()
    sd = synth.seeded_state_dict(ec, 4)
    for k in list(sd):
        if k.endswith(".1.2.weight"):
            sd[k] = -sd[k].abs()
    ec.load_state_dict(sd)
    x = torch.randn(64, 3, generator=torch.Generator().manual_seed(1))
    want = rignet_port.edge_conv_motion({"e." + k: v for k, v in sd.items()}, "e", data.pos, x, data.geo_edge_index)
    assert helpers.max_abs_diff(ec(data.pos, x, data.geo_edge_index), want) < 5e-6


def test_packed_weights_follow_load_state_dict(emulated):
    arch, kw, wseed, data, expect = helpers.load_golden("jointnet_b2_n256")
    model = helpers.build_model(arch, kw, wseed + 100)            # wrong weights first
    with torch.no_grad():
        wrong = model(data, data.pred_flow)
        assert helpers.max_abs_diff(wrong[2], expect[2]) > 1e-3
        model.load_state_dict(synth.seeded_state_dict(model, wseed))
        right = model(data, data.pred_flow)
    assert helpers.max_abs_diff(right[2], expect[2]) < 5e-6


def test_graph_cache_never_aliases_a_different_tensor(emulated):
    from morig_b200 import engine
    cache = engine.GraphCache()
    a = torch.tensor([[0, 1, 2], [1, 2, 0]])
    g1 = cache.get(a, 3)
    assert cache.get(a, 3) is g1                                  # same object, same version: hit
    b = a.clone()
    assert cache.get(b, 3) is not g1                              # equal content, different tensor: miss
    a[0, 0] = 2                                                   # in-place edit bumps the version: miss
    assert cache.get(a, 3) is not g1


def test_train_mode_and_cpu_tensors_fail_loudly():
    import morig_b200
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = morig_b200.jointnet_motion(**kw)
    data = synth.make_batch(1, 64, seed=0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.train()(data, data.pred_flow)                       # CPU tensors and no fallback, in either mode
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.eval()(data, data.pred_flow)
    mean_model = morig_b200.jointnet_motion(**dict(kw, aggr_method="mean")).train()
    assert mean_model.aggr_method == "mean"                       # (training with mean / max aggregation: see train_forward)


def test_factories_ignore_extra_kwargs_like_the_reference():
    import morig_b200
    m = morig_b200.jointnet_motion(num_keyframes=5, chn_output=3, aggr_method="attn", motion_dim=32, foo=1)
    assert isinstance(m, morig_b200.JointNetMotion)
    with pytest.raises(KeyError):
        morig_b200.skinnet_motion(num_keyframes=5)


def test_install_replaces_registry_entries():
    import types
    import morig_b200
    fake = types.ModuleType("models")
    fake.rignet = types.ModuleType("models.rignet")
    morig_b200.install(fake)
    assert fake.__dict__["jointnet_motion"] is morig_b200.jointnet_motion
    assert fake.rignet.SkinMotion is morig_b200.SkinMotion


def _declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            with open(os.path.join(ROOT, "include", fn)) as f:
                names += re.findall(r"MORIG_API[^;(]*?\b(morig_\w+)\s*\(", f.read())
    return sorted(set(names))


def test_c_abi_library_exports_every_declared_symbol():
    from morig_b200 import build
    build.build()                                                 # nvcc cross-compiles without a GPU
    declared = _declared_symbols()
    assert len(declared) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == declared                      # the ctypes table covers the whole header
    assert _lib.load().morig_version() == _lib.ABI_VERSION == 7


def test_tensor_core_weight_image_layout():
    """pack_tc_blob: TF32 hi/lo split + SWIZZLE_128B K-major stage image, checked with an independent
    address computation (byte offset = row*128 + ((chunk ^ (row % 8)) * 16) inside each [bn x 32] half)"""
    import numpy as np
    from morig_b200 import packing
    g = torch.Generator().manual_seed(0)
    n, k, bn = 300, 70, 256
    w = torch.randn(n, k, generator=g, dtype=torch.float64)
    blob, w_inv = packing.pack_tc_blob(w, 72, bn)
    blob = blob.numpy()
    assert w_inv == 1.0
    nk, nt = 3, 2
    assert blob.size == nt * nk * 2 * bn * 32
    img = blob.reshape(nt, nk, 2, bn * 32)
    rec = np.zeros((2, nt * bn, nk * 32), dtype=np.float32)
    for t in range(nt):
        for kc in range(nk):
            for half in range(2):
                flat = img[t, kc, half]
                for row in range(bn):
                    for c in range(8):
                        off = (row * 128 + ((c ^ (row % 8)) * 16)) // 4
                        rec[half, t * bn + row, kc * 32 + 4 * c: kc * 32 + 4 * c + 4] = flat[off:off + 4]
    hi, lo = rec
    assert not (hi.view(np.uint32) & 0x1FFF).any() and not (lo.view(np.uint32) & 0x1FFF).any()   # valid TF32 values
    full = np.zeros((nt * bn, nk * 32))
    full[:n, :k] = w.numpy()
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - full)
    assert err.max() <= 2.0 ** -21 * np.abs(full).max()
    assert np.abs(hi - full).max() <= 2.0 ** -11 * np.abs(full).max()


def test_tensor_core_weight_image_layout_fp16():
    """pack_tc_blob(kind=F16): power-of-two scaled fp16 hi/lo split in the same SWIZZLE_128B stage image, rows of
    64 halfs; hi + lo must reproduce W * 2^j to 2^-22 of the largest weight and nothing may overflow fp16"""
    import numpy as np
    from morig_b200 import packing
    g = torch.Generator().manual_seed(1)
    n, k, bn = 130, 100, 128
    w = torch.randn(n, k, generator=g, dtype=torch.float64) * 37.5
    blob, w_inv = packing.pack_tc_blob(w, 100, bn, packing.KIND_F16)
    blob = blob.numpy()
    nk, nt = 2, 2
    assert blob.dtype == np.float16 and blob.size == nt * nk * 2 * bn * 64
    img = blob.reshape(nt, nk, 2, bn * 64)
    rec = np.zeros((2, nt * bn, nk * 64), dtype=np.float64)
    for t in range(nt):
        for kc in range(nk):
            for half in range(2):
                flat = img[t, kc, half]
                for row in range(bn):
                    for c in range(8):
                        off = (row * 128 + ((c ^ (row % 8)) * 16)) // 2
                        rec[half, t * bn + row, kc * 64 + 8 * c: kc * 64 + 8 * c + 8] = flat[off:off + 8]
    assert np.isfinite(rec).all()
    full = np.zeros((nt * bn, nk * 64))
    full[:n, :k] = w.numpy()
    scaled_max = np.abs(full).max() / w_inv
    assert 2.0 ** 14 <= scaled_max < 2.0 ** 15
    assert np.log2(w_inv) == round(np.log2(w_inv))
    err = np.abs((rec[0] + rec[1]) * w_inv - full)
    assert err.max() <= 2.0 ** -22 * np.abs(full).max()


def test_amax_tracker_bookkeeping(emulated, monkeypatch):
    """AmaxTracker: slots are per storage, zeroed at the start of a forward, recomputed on demand for tensors
    this library did not write and after untracked in-place writes"""
    calls = []
    monkeypatch.setattr(engine, "_absmax", lambda t, off, ld, rows, cols, buf, idx: calls.append(("absmax", idx)))
    monkeypatch.setattr(engine, "_absmax_reset", lambda buf, idx: calls.append(("reset", idx)))
    ws = engine.Workspace()
    a, b = torch.zeros(4, 8), torch.zeros(4, 8)
    with engine.forward_scope(ws, "cpu"):
        tr = engine._tracker
        assert tr is ws.tracker
        p_a = engine._amax_in(a, 0, 8, 4, 8)                  # foreign tensor: computed on demand, once
        assert engine._amax_in(a[1:], 8, 8, 3, 8) == p_a      # a view shares the storage slot
        assert calls == [("absmax", 0)]
        p_b = engine._amax_out(b)                             # tracked producer: no absmax launch needed
        assert p_b == p_a + 4 and engine._amax_in(b, 0, 8, 4, 8) == p_b and len(calls) == 1
        engine._untracked(b)                                  # e.g. row_normalize in place
        assert calls[-1] == ("reset", 1)
        engine._amax_in(b, 0, 8, 4, 8)
        assert calls[-1] == ("absmax", 1)
        with engine.forward_scope(ws, "cpu"):                 # nested scope: no reset
            assert engine._tracker is tr and 0 in tr.valid
    assert engine._tracker is None
    with engine.forward_scope(ws, "cpu"):
        assert not ws.tracker.valid                           # new forward: everything recomputed / re-tracked


def test_c_abi_argument_validation_without_gpu():
    """entry points must reject bad descriptors with a code + message instead of launching"""
    lib = _lib.load()
    d = _lib.DenseDesc()
    assert lib.morig_dense_fwd(ctypes.byref(d), None) == 1001
    assert b"null operand" in lib.morig_last_error()
    e = _lib.EdgeDesc()
    assert lib.morig_edgeconv_fwd(ctypes.byref(e), None) == 1001
    assert lib.morig_graph_prep(None, 10, 0, None, None, None, None, 0, None) == 1001


def test_packed_weights_follow_in_place_parameter_edits(emulated):
    """stale-pack protection (weights_fingerprint): optimizer-style in-place updates, nn.init and load_state_dict on a
    plain nn.Sequential child / nested module must all be seen by the next forward"""
    from oracle import rignet_port
    import morig_b200
    data = synth.make_batch(1, 100, seed=5)
    x = torch.randn(100, 64, generator=torch.Generator().manual_seed(0))
    rig = morig_b200.GCNRig(64, 3).eval()
    rig.load_state_dict(synth.seeded_state_dict(rig, 2))

    def check():
        sd = {"r." + k: v for k, v in rig.state_dict().items()}
        want = rignet_port.gcn_rig(sd, "r", data.pos, x, data.tpl_edge_index, data.geo_edge_index, data.batch)
        got = rig(data.pos, x, data.tpl_edge_index, data.geo_edge_index, data.batch)
        assert helpers.max_abs_diff(got, want) < 5e-6 * max(1.0, float(want.abs().max()))
        return got

    a = check()
    with torch.no_grad():
        for p in rig.parameters():
            p.add_(0.01 * torch.ones_like(p))                                  # what an optimizer step does
    b = check()
    assert not torch.equal(a, b)
    torch.nn.init.constant_(rig.mlp_transform[1].bias, 0.5)
    check()
    child = rig.mlp_glb                                                         # plain nn.Sequential, no hooks of ours
    child.load_state_dict({k: v * 1.5 for k, v in child.state_dict().items()})
    check()
    nested = rig.gcu_2                                                          # the pack lives on the parent GCNRig
    nested.load_state_dict({k: (v * 0.5 if v.is_floating_point() else v) for k, v in nested.state_dict().items()})
    check()


def test_standalone_forwards_validate_shapes(emulated):
    import morig_b200
    data = synth.make_batch(1, 64, seed=1)
    rig = morig_b200.GCNRig(64, 3).eval()
    good = torch.zeros(64, 64)
    with pytest.raises(ValueError):
        rig(data.pos, torch.zeros(64, 32), data.tpl_edge_index, data.geo_edge_index, data.batch)
    with pytest.raises(ValueError):
        rig(data.pos[:, :2].contiguous(), good, data.tpl_edge_index, data.geo_edge_index, data.batch)
    with pytest.raises(ValueError):
        rig(data.pos, good, data.tpl_edge_index, data.geo_edge_index, data.batch[:10])
    gcu = morig_b200.GCUMotion(64, 256).eval()
    with pytest.raises(ValueError):
        gcu(data.pos, torch.zeros(64, 60), data.tpl_edge_index, data.geo_edge_index)
    ec = morig_b200.EdgeConvMotion(morig_b200.MLP([6, 32, 32]), morig_b200.MLP([6, 16, 16])).eval()
    with pytest.raises(ValueError):
        ec(torch.zeros(64, 4), torch.zeros(64, 3), data.tpl_edge_index)


def test_operand_range_tags_are_only_trusted_for_the_exact_tensor_state():
    """train_ops tags a produced matrix with the device scalar holding max |value| (reduced by the producing kernel); a
    consumer may use it only for that tensor object in that state -- views, copies and in-place edits fall back to the
    |max| pass"""
    from morig_b200 import train_ops as T
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    a = T._new_amax(t.device)
    b = T._new_amax(t.device)
    assert a.data_ptr() != b.data_ptr() and float(a) == 0.0 and a.numel() == 1       # slices of one zeroed block
    assert T._known_amax(t) == 0
    T._tag_amax(t, a)
    assert T._known_amax(t) == a.data_ptr()
    assert T._known_amax(t[:, :2]) == 0 and T._known_amax(t.clone()) == 0            # a view / a copy: not tagged
    t.add_(1.0)                                                                      # in-place edit: version moved on
    assert T._known_amax(t) == 0


def test_tile_stream_arithmetic_of_the_persistent_kernels(tmp_path):
    """morig_b200/csrc/tile_iter.cuh (the header the tcgen05 kernels include): the incremental tile iterator visits exactly
    the coordinates of TileMap::decode, for 20 000 random tile grids / CTA strides; host build with g++"""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "tile_iter_test"
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "tile_iter_test.cpp")
    subprocess.run([cxx, "-O1", "-std=c++17", "-o", str(exe), src], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout


@pytest.mark.parametrize("rbw,tile_rows", [(2, 128), (4, 256)])
def test_segmax_epilogue_protocol_model(rbw, tile_rows):
    """Executable model of the fused-EdgeConv epilogue protocol (csrc/gemm_tc.cuh epilogue_segmax_role) and of its contract with
    morig_fill_cut_f32: every epilogue warp owns `rbw` contiguous 32-row blocks of a tile, carries a segment that runs across
    its blocks, stores a segment it sees completely with a PLAIN store and merges one that leaves its range with an atomic
    max on a destination that was pre-filled with -inf iff the segment straddles a multiple of 32 slots.  The model replays
    those rules in a random order of warps on random segment layouts (lengths 1..700, ragged ends) over an output that
    starts as garbage, and checks (1) the result is the segmented max, (2) no vertex receives a plain store AND any other
    write (the race the selective fill must exclude), (3) atomics only ever hit pre-filled vertices."""
    import numpy as np
    rng = np.random.default_rng(rbw)
    for trial in range(60):
        lens = rng.choice([1, 2, 7, 17, 31, 32, 33, 63, 64, 65, 127, 129, 255, 257, 700], size=rng.integers(3, 60))
        tgt = np.repeat(np.arange(len(lens)), lens)                      # CSR target of every slot (sorted)
        M = len(tgt)
        val = rng.standard_normal(M)
        rowptr = np.concatenate([[0], np.cumsum(lens)])
        a, b = rowptr[:-1], rowptr[1:]
        prefilled = (a >> 5) != ((b - 1) >> 5)                           # morig_fill_cut_f32
        out = np.full(len(lens), 123.0)                                  # garbage start values
        out[prefilled] = -np.inf
        plain = np.zeros(len(lens), int)
        atomic = np.zeros(len(lens), int)
        span = 32 * rbw
        ranges = list(range(0, (M + tile_rows - 1) // tile_rows * tile_rows, span))
        rng.shuffle(ranges)                                              # warps / CTAs finish in any order
        key = lambda r: tgt[r] if 0 <= r < M else (-2 if r < 0 else -1)
        for r0 in ranges:
            carry, open_cut = None, False
            for i in range(rbw):
                base = r0 + 32 * i
                k = [key(base + j) for j in range(32)]
                first_cut = k[0] == key(base - 1)
                last_cut = k[31] == key(base + 32)
                tails = [j for j in range(32) if j == 31 or k[j] != k[j + 1]]
                cont = i > 0 and first_cut
                if i == 0:
                    open_cut = first_cut
                defer = last_cut and i < rbw - 1
                run, w = None, []
                for j in range(32):
                    v = val[base + j] if base + j < M else 0.0
                    head = (j == 0 and not cont) or (j > 0 and k[j] != k[j - 1])
                    run = v if head else max(run if run is not None else carry, v)
                    w.append(run)
                for j in tails:
                    if defer and j == 31:
                        continue
                    if k[j] < 0:
                        continue
                    at = (j == tails[0] and first_cut and open_cut) or (j == 31 and last_cut)
                    if at:
                        assert prefilled[k[j]], "atomic merge into a vertex that was not pre-filled"
                        out[k[j]] = max(out[k[j]], w[j])
                        atomic[k[j]] += 1
                    else:
                        out[k[j]] = w[j]
                        plain[k[j]] += 1
                carry = w[31]
                open_cut = defer and tails[0] == 31 and first_cut and open_cut
        want = np.array([val[s:e].max() for s, e in zip(a, b)])
        assert np.array_equal(out, want)
        assert ((plain == 1) & (atomic == 0) | (plain == 0) & (atomic >= 1)).all()
