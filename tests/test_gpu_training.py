"""GPU tier of the training path: the real kernels of csrc/train.cu through the C-ABI against
(1) the torch-CPU restatements of each entry point (tests/emu.py), (2) the fp64 training oracle at module level, and
(3) whole networks forward + backward; plus the integer-exact argmax KAT and an optimisation-step smoke test."""
import pytest
import torch

import emu
import helpers
import test_training as tt
from morig_b200 import autograd_ops as A
from morig_b200 import engine, synth, train_ops as T
from oracle import rignet_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def g_(seed):
    return torch.Generator().manual_seed(seed)


def close(a, b, tol=2e-5):
    b = b.double()
    return float((a.detach().cpu().double() - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


# ---- entry points against their restatements ----------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(1, 32, 128), (130, 36, 64), (1000, 3, 64), (4099, 544, 512), (257, 838, 1024), (20, 1024, 7),
                                   (777, 66, 128)])
def test_linear_fwd_input_grad_and_wgrad(M, K, N):
    g = g_(M + K + N)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    for relu in (False, True):
        assert close(T.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV), relu), emu._t_linear_fwd(x.double(), w.double(), b.double(), relu))
    assert close(T.matmul_nn(dy.to(DEV), w.to(DEV)), dy.double() @ w.double())
    dW, db = T.wgrad(dy.to(DEV), x.to(DEV), True)
    assert close(dW, dy.double().t() @ x.double(), 1e-5) and close(db, dy.double().sum(0), 1e-5)


@pytest.mark.parametrize("M,N,K,pad", [(5000, 64, 64, 0), (40000, 200, 130, 3), (3000, 1024, 300, 0), (2048, 128, 1000, 5),
                                       (70001, 256, 256, 0)])
def test_wgrad_tensor_core_kernel(M, N, K, pad):
    """csrc/wgrad_tc.cuh (3xTF32 tcgen05, transposing producers): ragged tile edges, row slices that do not fill a stage,
    operands that are column slices of wider buffers; and an integer-valued case that must be exact"""
    g = g_(M + N + K)
    dy_full = torch.randn(M, N + pad, generator=g).to(DEV)
    x_full = (torch.randn(M, K + 2 * pad, generator=g) * 3 + 0.5).to(DEV)
    dy, x = dy_full[:, pad:], x_full[:, pad:pad + K]
    dW, db = T.wgrad(dy, x, True)
    assert close(dW, dy.cpu().double().t() @ x.cpu().double(), 1e-5) and close(db, dy.cpu().double().sum(0), 1e-5)
    dyi = torch.randint(-3, 4, (M, N), generator=g).float().to(DEV)
    xi = torch.randint(-4, 5, (M, K), generator=g).float().to(DEV)
    dWi, dbi = T.wgrad(dyi, xi, True)
    assert torch.equal(dWi.cpu().double(), dyi.cpu().double().t() @ xi.cpu().double())
    assert torch.equal(dbi.cpu().double(), dyi.cpu().double().sum(0))


def test_kernels_take_column_slices_in_place():
    """row strides are passed down: column slices of wider buffers (P / Q halves, key-frame slices of the flow, slices
    of a concatenated gradient) are used without copies"""
    g = g_(0)
    big = torch.randn(500, 40, generator=g).to(DEV)
    w = torch.randn(16, 12, generator=g).to(DEV)
    xs = big[:, 7:19]
    assert xs.stride(0) == 40
    assert close(T.linear_fwd(xs, w, None), xs.cpu().double() @ w.cpu().double().t())
    dy = torch.randn(500, 60, generator=g).to(DEV)[:, 3:19]
    dW, _ = T.wgrad(dy, xs, False)
    assert close(dW, dy.cpu().double().t() @ xs.cpu().double(), 1e-5)
    assert close(T.concat_cols([xs, dy]), torch.cat([xs.cpu(), dy.cpu()], 1), 0)


@pytest.mark.parametrize("R,C", [(2, 8), (50, 16), (4096, 64), (100000, 256), (7, 1024)])
def test_batchnorm_train_forward_and_backward(R, C):
    g = g_(R + C)
    x = torch.relu(torch.randn(R, C, generator=g) * 2 + 0.3)
    x[:, 0] = 0.0                                                       # a dead channel: variance 0
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.5
    dy = torch.randn(R, C, generator=g)
    rm_e, rv_e = rm.clone().double(), rv.clone().double()
    y_e, mu_e, is_e = emu._t_bn_train_fwd(x.double(), gamma.double(), beta.double(), rm_e, rv_e, 0.1)
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    y, mu, istd = T.bn_train_fwd(x.to(DEV), gamma.to(DEV), beta.to(DEV), rm_d, rv_d, 0.1)
    assert close(y, y_e) and close(mu, mu_e) and close(istd, is_e) and close(rm_d, rm_e) and close(rv_d, rv_e)
    for relu in (True, False):
        dz_e, dg_e, db_e = emu._t_bn_relu_bwd(dy.double(), x.double(), gamma.double(), mu_e, is_e, relu)
        dz, dg, db = T.bn_relu_bwd(dy.to(DEV), x.to(DEV), gamma.to(DEV), mu, istd, relu)
        assert close(dz, dz_e, 1e-4) and close(dg, dg_e, 1e-4) and close(db, db_e, 1e-5)


def _graph(n, e, seed):
    g = g_(seed)
    ei = torch.randint(0, max(n - 3, 1), (2, e), generator=g)
    if e > 100:
        ei[1, : e // 5] = 2                                             # one heavy target
    gr = engine.graph_prep(ei.to(DEV), n)
    gr.e_real = int(gr.rowptr[n].item())
    cpu = engine.Graph(rowptr=gr.rowptr.cpu(), col=gr.col.cpu(), tgt=gr.tgt.cpu(), n=n, e_max=gr.e_max)
    cpu.e_real = gr.e_real
    return gr, cpu


@pytest.mark.parametrize("n,e,C", [(5, 3, 16), (300, 2500, 32), (1500, 12000, 128)])
def test_edge_gather_relu_forward_and_backward(n, e, C):
    gr, cg = _graph(n, e, n + C)
    g = g_(C)
    pq = torch.randn(n, 2 * C, generator=g)
    h = T.edge_gather_relu(pq[:, :C].to(DEV), pq[:, C:].to(DEV), gr)
    h_e = emu._t_edge_gather_relu(pq[:, :C], pq[:, C:], cg)
    assert torch.equal(h.cpu(), h_e)
    dh = torch.randn(gr.e_real, C, generator=g)
    assert close(T.edge_gather_relu_bwd(dh.to(DEV), h, gr), emu._t_edge_gather_relu_bwd(dh.double(), h_e.double(), cg), 1e-5)


@pytest.mark.parametrize("n,e,C", [(5, 3, 16), (300, 2500, 33), (1500, 12000, 128)])
def test_segmax_with_ties_follows_first_row_rule(n, e, C):
    gr, cg = _graph(n, e, n + C + 1)
    g = g_(C + 1)
    y = torch.randint(-3, 4, (gr.e_real, C), generator=g).float()        # small integers: many exact ties
    out, arg = T.segmax_fwd(y.to(DEV), gr.rowptr, n)
    out_e, arg_e = emu._t_segmax_fwd(y, cg.rowptr, n)
    assert torch.equal(out.cpu(), out_e) and torch.equal(arg.cpu(), arg_e)
    dout = torch.randn(n, C, generator=g)
    assert torch.equal(T.segmax_bwd(dout.to(DEV), arg, gr.e_real).cpu(), emu._t_segmax_bwd(dout, arg_e, gr.e_real))


def test_graph_pooling_pieces():
    g = g_(3)
    batch = torch.sort(torch.randint(0, 5, (3000,), generator=g)).values
    batch[0], batch[-1] = 0, 4
    b32 = batch.to(torch.int32).to(DEV)
    ptr = T.seg_ptr(b32, 5)
    assert torch.equal(ptr.cpu(), emu._t_seg_ptr(batch.to(torch.int32), 5))
    x = torch.randn(3000, 100, generator=g)
    out, arg = T.segmax_fwd(x.to(DEV), ptr, 5)
    out_e, arg_e = emu._t_segmax_fwd(x, ptr.cpu(), 5)
    assert torch.equal(out.cpu(), out_e) and torch.equal(arg.cpu(), arg_e)
    xg = torch.randn(5, 64, generator=g)
    assert torch.equal(T.row_gather(xg.to(DEV), b32).cpu(), xg[batch])
    assert close(T.seg_sum(x.to(DEV), ptr, 5), emu._t_seg_sum(x.double(), ptr.cpu(), 5), 1e-6)


def test_normalize_and_attention_kernels():
    g = g_(4)
    x = torch.randn(1000, 32, generator=g)
    x[3] = 0.0
    dy = torch.randn(1000, 32, generator=g)
    assert close(T.normalize_fwd(x.to(DEV)), emu._t_normalize_fwd(x.double()), 1e-6)
    ok = torch.ones(1000, dtype=torch.bool); ok[3] = False
    assert close(T.normalize_bwd(x.to(DEV), dy.to(DEV))[ok], emu._t_normalize_bwd(x.double(), dy.double())[ok], 1e-5)
    N, Tn, HD, d = 777, 5, 128, 64
    q0, kc, vc = (torch.randn(HD, generator=g) for _ in range(3))
    Kx, Vx = torch.randn(N, Tn, HD, generator=g), torch.randn(N, Tn, HD, generator=g)
    dout = torch.randn(N, HD, generator=g)
    out, att = T.attn_cls_fwd(*(t.to(DEV) for t in (q0, kc, vc, Kx, Vx)), d)
    out_e, att_e = emu._t_attn_cls_fwd(*(t.double() for t in (q0, kc, vc, Kx, Vx)), d)
    assert close(out, out_e, 1e-5) and close(att, att_e, 1e-5)
    got = T.attn_cls_bwd(*(t.to(DEV) for t in (q0, kc, vc, Kx, Vx)), att, dout.to(DEV), d)
    want = emu._t_attn_cls_bwd(*(t.double() for t in (q0, kc, vc, Kx, Vx)), att_e, dout.double(), d)
    for a, b in zip(got, want):
        assert close(a, b, 2e-5)


# ---- modules: train-mode forward + backward against the fp64 training oracle ---------------------------------------------
def _module_case(mod, oracle_fn, call, inputs, seed):
    """returns worst errors (outputs, parameter gradients) of the CUDA path and of the fp32 oracle w.r.t. fp64"""
    mod.load_state_dict(synth.seeded_state_dict(mod, seed))
    sd0 = {"m." + k: v.clone() for k, v in mod.state_dict().items()}

    def run_oracle(dtype):
        sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
        ps = [k for k in sd if sd[k].is_floating_point() and not k.endswith(("running_mean", "running_var"))]
        for k in ps:
            sd[k].requires_grad_(True)
        ins = [(t.to(dtype) if t.is_floating_point() else t) for t in inputs]
        with rignet_port.training_mode():
            out = oracle_fn(sd, *ins)
        w = torch.linspace(-1, 1, out.numel(), dtype=dtype).reshape(out.shape)
        (out * w).sum().backward()
        return out.detach(), {k[2:]: sd[k].grad for k in ps}

    o32, o64 = run_oracle(torch.float32), run_oracle(torch.float64)
    mod = mod.to(DEV).train()
    out = call(mod, *[t.to(DEV) for t in inputs])
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape).to(DEV)
    (out * w).sum().backward()
    ok, err = tt.close_to_truth(out, o32[0], o64[0], tol=2e-5)
    assert ok, ("output", err)
    for k, p in mod.named_parameters():
        ok, err = tt.close_to_truth(p.grad, o32[1][k], o64[1][k], tol=1e-4)
        assert ok, (k, err, float(o64[1][k].abs().max()))


def test_train_mode_modules_match_oracle():
    import morig_b200
    data = synth.make_batch(2, 400, seed=3)
    g = g_(1)
    x = torch.randn(800, 64, generator=g)
    _module_case(morig_b200.EdgeConvMotion(morig_b200.MLP([128, 128, 128]), morig_b200.MLP([6, 16, 16])),
                 lambda sd, pos, x, ei: rignet_port.edge_conv_motion(sd, "m", pos, x, ei), lambda m, pos, x, ei: m(pos, x, ei),
                 [data.pos, x, data.geo_edge_index], 7)
    _module_case(morig_b200.GCUMotion(64, 256),
                 lambda sd, pos, x, a, b: rignet_port.gcu_motion(sd, "m", pos, x, a, b), lambda m, pos, x, a, b: m(pos, x, a, b),
                 [data.pos, x, data.tpl_edge_index, data.geo_edge_index], 8)
    _module_case(morig_b200.GCU(64, 128),
                 lambda sd, x, a, b: rignet_port.gcu(sd, "m", x, a, b), lambda m, x, a, b: m(x, a, b),
                 [x, data.tpl_edge_index, data.geo_edge_index], 5)
    xs = torch.nn.functional.normalize(torch.randn(600, 5, 32, generator=g), dim=2)
    _module_case(morig_b200.TemporalAttn(32, 2, 64, 512, 64),
                 lambda sd, x: rignet_port.temporal_attn(sd, "m", x), lambda m, x: m(x), [xs], 3)


@pytest.mark.parametrize("arch,b,n", [("jointnet_motion", 2, 256), ("jointnet_motion", 1, 1024), ("skinnet_motion", 2, 256)])
def test_whole_network_training_step_matches_oracle(arch, b, n):
    """forward + backward of a whole network in train mode: outputs within 1e-4 (or a small multiple of the fp32
    reference's own distance from the fp64 result), gradients by relative L2 norm over all parameters, running
    statistics, batch counters"""
    kw = synth.ARCH_KWARGS[arch]
    model = helpers.build_model(arch, kw, 3)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    data = synth.make_batch(b, n, seed=21, with_skin=(arch == "skinnet_motion"))
    o_outs, o_grads, o_bufs = tt.oracle_train_step(arch, kw, sd0, data, data.pred_flow)
    x_outs, x_grads, x_bufs = tt.oracle_train_step(arch, kw, sd0, data, data.pred_flow, dtype=torch.float64)
    model = model.to(DEV).train()
    dd = data.to(DEV)
    outs = model(dd, dd.pred_flow)
    for a, e, x, k in zip(outs, o_outs, x_outs, helpers.OUT_KEYS):
        ok, err = tt.close_to_truth(a, e, x)
        assert ok, (k, err)
    tt.loss_of(outs, dd.pos).backward()
    torch.cuda.synchronize()
    ours, ref = tt.grad_l2_errors(dict(model.named_parameters()), o_grads, x_grads)
    print(f"{arch} {b}x{n}: relative L2 gradient error vs fp64: cuda {ours:.3e}, fp32 oracle {ref:.3e}")
    assert ours < max(1e-3, 10.0 * ref), (ours, ref)
    for k, v in model.state_dict().items():
        if k in o_bufs:
            ok, err = tt.close_to_truth(v, o_bufs[k], x_bufs[k], tol=1e-5)
            assert ok, (k, err)
        if k.endswith("num_batches_tracked"):
            assert int(v) == (5 if k.startswith("motionNet") else 1), k


def test_integer_exact_kat_argmax_and_gradients():
    """integer per-edge values with deliberate ties (duplicate edges): maxima, argmax slots and the
    routed gradients must equal the oracle's first-edge rule bit for bit"""
    g = g_(0)
    n = 500
    ei = torch.randint(0, n, (2, 4000), generator=g)
    ei = torch.cat([ei, ei[:, :1500]], dim=1)                            # duplicate edges: exact ties
    gr = engine.graph_prep(ei.to(DEV), n)
    gr.e_real = int(gr.rowptr[n].item())
    x = torch.randint(-2, 3, (n, 48), generator=g).float()
    xd = x.to(DEV).requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    tgt, col = gr.tgt[:gr.e_real].long(), gr.col[:gr.e_real].long()
    y = xd[tgt] * 2 + xd[col]                                            # integer per-edge values
    out, arg = A.SegMax.apply(y, gr.rowptr, n)
    (out * torch.arange(48, device=DEV)).sum().backward()
    ref = rignet_port.normalized_edges(ei, n)
    yo = xo[ref[1]] * 2 + xo[ref[0]]
    out_o = rignet_port._segment_max(yo, ref[1], n)
    (out_o * torch.arange(48)).sum().backward()
    assert torch.equal(out.detach().cpu(), out_o.detach())
    assert torch.equal(xd.grad.cpu(), xo.grad)
    # argmax slot -> (target, source) must be the oracle's first maximal edge of that target
    a = arg.cpu().long()
    src_of_arg = gr.col.cpu().long()[a]
    first = torch.full((n, 48), ref.shape[1], dtype=torch.long)
    hit = yo.detach() == out_o.detach()[ref[1]]
    rows = torch.arange(ref.shape[1]).unsqueeze(1).expand_as(hit)
    first = first.scatter_reduce(0, ref[1].unsqueeze(1).expand_as(hit), torch.where(hit, rows, torch.full_like(rows, ref.shape[1])),
                                 reduce="amin", include_self=True)
    assert torch.equal(src_of_arg, ref[0][first])


def test_optimisation_steps_and_eval_after_training():
    """a few SGD steps of the reference's training recipe shape (training/train_rig.py:186-191): the loss goes down, and
    eval() afterwards runs the fused inference kernels on the UPDATED weights and running statistics.  Train-mode BatchNorm
    on seeded random weights gives a gradient norm of ~280 at a loss of 0.2, so plain SGD only descends monotonically for
    learning rates <= 1e-6 (scripts/train_probe.py: the fp32 and the tensor-core GEMM paths then agree to 4 digits per
    step); larger rates wander for either engine"""
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 3, DEV).train()
    data = synth.make_batch(2, 256, seed=5).to(DEV)
    target = torch.tanh(torch.randn(512, 3, generator=g_(9))).to(DEV) * 0.1
    opt = torch.optim.SGD(model.parameters(), lr=1e-7)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        _, _, pred = model(data, data.pred_flow)
        loss = (torch.tanh(pred) - target).pow(2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(torch.isfinite(torch.tensor(losses))) and all(b < a for a, b in zip(losses, losses[1:])), losses
    model.eval()
    with torch.no_grad():
        out = model(data, data.pred_flow)
    expect = helpers.oracle_forward("jointnet_motion", kw, model, data.to("cpu"), data.pred_flow.cpu())
    for o, e, k in zip(out, expect, helpers.OUT_KEYS):
        assert helpers.max_abs_diff(o, e) < helpers.TOL * max(1.0, float(e.abs().max())), k


@pytest.mark.parametrize("N,K,bn", [(130, 100, 128), (256, 256, 256), (1024, 838 // 2 * 2, 256), (40, 64, 128)])
def test_device_side_weight_image_equals_host_packer(N, K, bn):
    """morig_pack_tc_f16 (weights change every training step, so the fp16-split tensor-core image is built on the device)
    against packing.pack_tc_blob: the same bytes, for the weight and for its transpose (input-gradient GEMM)"""
    from morig_b200 import _lib, packing
    lib = _lib.load()
    w = torch.randn(N, K, generator=g_(N + K))
    for transposed in (False, True):
        logical = w.t().contiguous() if transposed else w                      # the operand [n_out, k] the GEMM multiplies by
        n_out, k = logical.shape
        want, w_inv = packing.pack_tc_blob(logical.double(), k, bn, packing.KIND_F16)
        blob = torch.zeros(lib.morig_pack_tc_f16_bytes(n_out, k, bn), dtype=torch.uint8, device=DEV)
        scal = torch.zeros(2, device=DEV)
        wd = w.to(DEV)
        _lib.check(lib.morig_pack_tc_f16(wd.data_ptr(), K, n_out, k, 1 if transposed else 0, bn, blob.data_ptr(), scal.data_ptr(),
                                         scal.data_ptr() + 4, _lib.stream_ptr()), "pack")
        assert float(scal[0]) == w_inv
        assert torch.equal(blob.cpu().view(torch.float16), want.view(torch.float16))


@pytest.mark.parametrize("M,K,N", [(4099, 544, 512), (1000, 64, 128), (300000, 256, 256)])
def test_training_gemms_on_the_tensor_core_engine(M, K, N):
    """forward and input-gradient GEMMs of the training path take the tcgen05 split-fp16 engine for large shapes:
    fp32-class accuracy against fp64"""
    g = g_(M + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    assert T._tc_ok(M, K, N, x.to(DEV))
    got = T.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV), True)
    assert close(got, torch.relu(x.double() @ w.double().t() + b.double()), 1e-5)
    assert close(T.matmul_nn(dy.to(DEV), w.to(DEV)), dy.double() @ w.double(), 1e-5)
