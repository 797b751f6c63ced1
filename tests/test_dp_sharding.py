"""Multi-process (gloo, world_size 2, CPU) test of the data-parallel host logic: meshes partitioned across
ranks, each rank runs its shard (CUDA entry points emulated on CPU), outputs gathered in rank order equal
the single-process batch; timing reduction is a max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morig_b200 import dp, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    import emu
    import helpers

    class MP:
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    emu.install(MP())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sizes = [100, 64, 144, 81, 100]
    meshes = [synth.make_mesh(n, 500 + i) for i, n in enumerate(sizes)]
    bins = dp.partition([dp.mesh_cost(m) for m in meshes], world)
    mine = bins[rank]
    kw = synth.ARCH_KWARGS["jointnet_motion"]
    model = helpers.build_model("jointnet_motion", kw, 21)
    with torch.no_grad():
        local = synth.collate([meshes[i] for i in mine])
        out = model(local, local.pred_flow)[2]
    rows = [sum(sizes[i] for i in b) for b in bins]
    gathered = dp.gather_rows(out, rows)
    slowest = dp.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        with torch.no_grad():
            order = [i for b in bins for i in b]
            full = synth.collate([meshes[i] for i in order])
            ref = model(full, full.pred_flow)[2]
        ret["err"] = float((gathered - ref).abs().max())
        ret["slowest"] = slowest
        ret["bins"] = bins
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_batch():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["err"] < 1e-6
        assert ret["slowest"] == 2.0
        bins = ret["bins"]
    assert sorted(i for b in bins for i in b) == [0, 1, 2, 3, 4]


def test_partition_balances_and_is_deterministic():
    costs = [23 * n for n in (4096, 1024, 2048, 4096, 512, 1024, 3000, 800)]
    bins = dp.partition(costs, 4)
    assert bins == dp.partition(costs, 4)
    loads = [sum(costs[i] for i in b) for b in bins]
    assert max(loads) <= 1.35 * (sum(costs) / 4)
    assert dp.partition([5, 5], 4) == [[0], [1], [], []]


def _train_worker(rank, world, port, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.dirname(here))
    import emu
    import helpers
    import test_training as tt

    class MP:
        def setattr(self, obj, name, val):
            setattr(obj, name, val)
    emu.install(MP()); emu.install_train(MP())
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    kw = synth.ARCH_KWARGS["masknet_motion"]
    model = helpers.build_model("masknet_motion", kw, 4).train()
    ar = dp.GradAllReduce(model, bucket_mb=8.0)
    assert len(ar.buckets) >= 2
    shards = [synth.make_batch(1, 64, seed=300 + r) for r in range(world)]
    ar.zero_grad()
    tt.loss_of(model(shards[rank], shards[rank].pred_flow), shards[rank].pos).backward()
    nbytes = ar.finish()
    got = {k: p.grad.clone() for k, p in model.named_parameters()}
    flat_is_grad = all(p.grad.data_ptr() >= ar.flat.data_ptr() for p in model.parameters())
    if rank == 0:
        # single-process reference: mean over the two shards of the per-shard gradients (fresh copies of the model:
        # per-replica BatchNorm statistics, as under DistributedDataParallel)
        want = None
        for r in range(world):
            m = helpers.build_model("masknet_motion", kw, 4).train()
            tt.loss_of(m(shards[r], shards[r].pred_flow), shards[r].pos).backward()
            g = {k: p.grad for k, p in m.named_parameters()}
            want = g if want is None else {k: want[k] + g[k] for k in g}
        ret["err"] = max(float((got[k] - want[k] / world).abs().max()) / max(1.0, float(want[k].abs().max())) for k in got)
        ret["bytes"] = nbytes
        ret["numel"] = sum(p.numel() for p in model.parameters())
        ret["views"] = flat_is_grad
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_equals_mean_of_shard_gradients():
    """dp.GradAllReduce under gloo: flat-buffer, bucketed all-reduce launched from post-accumulate hooks during the
    backward; the result is the mean of the per-rank gradients and the whole parameter set travels once"""
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_train_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret["err"] < 1e-6
        assert ret["bytes"] == 4 * ret["numel"] and ret["views"]
