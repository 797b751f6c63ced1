"""Strong-scaling line on RAGGED meshes (BASELINE.json configs[4] as real data would look: 64 meshes of 1K-5K vertices):
the fixed set is split over the ranks by `dp.partition` (greedy by edge count), every rank runs its shard in batches of at
most 8 meshes through jointnet_motion, time = max over ranks (CUDA events), value = 64 meshes / time.

    python scripts/scale_ragged.py                                              # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/scale_ragged.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import morig_b200
from morig_b200 import dp, synth

SIZES = [1024, 1600, 2048, 2500, 3072, 3600, 4096, 4900]          # torus grids; 8 of each = 64 meshes
ARCH = "jointnet_motion"


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    meshes = [synth.make_mesh(SIZES[i % len(SIZES)], 900 + i) for i in range(64)]
    bins = dp.partition([dp.mesh_cost(m) for m in meshes], world)
    mine = [meshes[i] for i in bins[rank]]
    batches = [synth.collate(mine[i:i + 8]).to(dev) for i in range(0, len(mine), 8)]
    kw = synth.ARCH_KWARGS[ARCH]
    model = getattr(morig_b200, ARCH)(**kw).eval()
    model.load_state_dict(synth.seeded_state_dict(model, 1))
    model = model.to(dev)

    def run():
        with torch.no_grad():
            for b in batches:
                model(b, b.pred_flow)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    ms = dp.max_over_ranks(s.elapsed_time(e) / reps, dev)
    loads = [sum(dp.mesh_cost(meshes[i]) for i in b) for b in bins]
    if rank == 0:
        print(json.dumps({"metric": "meshes/sec jointnet_motion fwd, 64 ragged meshes (1K-5K vertices), strong scaling",
                          "value": 64 / (ms / 1e3), "unit": "meshes/s", "n_gpus": world, "ms_per_pass": ms, "scaling": "strong",
                          "meshes_per_rank": [len(b) for b in bins], "load_imbalance": max(loads) / (sum(loads) / world),
                          "vertices_total": sum(m["pos"].shape[0] for m in meshes)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
