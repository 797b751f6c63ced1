#!/usr/bin/env python
"""Benchmark of the hot path: meshes/sec of the `jointnet_motion` forward on synthetic 4096-vertex
meshes (BASELINE.json `metric`, configs[1]: batch = 4 meshes per GPU), fp32, eval mode.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]           # reference CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU, weak scaling

A "step" is one forward over one batch of 4 meshes per GPU (graph preparation included: the reference
feeds a new Batch every iteration, training/train_rig.py:206-223).  Prints ONE JSON line on rank 0.

  value   meshes/s, inputs resident in HBM, each step timed with CUDA events on the launch stream,
          an L2 flush (256 MiB write) between steps, max over ranks
  e2e     the same metric through the public API with HOST (pinned) inputs and HOST results: every step's H2D copy
          of the batch, forward and D2H copy of the three outputs are inside the timed region.  Steps go through
          `morig_b200.HostPipeline` (copy-in / forward / copy-out of consecutive batches on three streams);
          `e2e.serial` is the strictly sequential `model(data.to(dev), flow)` + `.to("cpu")` loop of the reference
  roofline / cpu_baseline: see DESIGN.md "Measurement"
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

ARCH = "jointnet_motion"
MESHES_PER_GPU = 4
N_VTX = 4096
METRIC = "meshes/sec jointnet_motion fwd, 4K-vtx synthetic"
UNIT = "meshes/s"
# canonical algorithmic FLOPs of the reference formulation, SURVEY.md §8(d): 105.25 MFLOP per vertex
ALG_FLOP_PER_VERTEX = 105.25e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def workload_config(world: int, meshes_per_gpu: int) -> dict:
    """`config` of the bench line -- the same object in both arms (ours and `--impl reference`)"""
    return {"workload": f"{ARCH} forward, batch={meshes_per_gpu} x {N_VTX}-vertex synthetic meshes per GPU "
                        "(BASELINE.json configs[1]), graph preparation included every step",
            "arch": ARCH, "n_vtx": N_VTX, "meshes_per_gpu": meshes_per_gpu,
            "l2": "256 MiB flush write between timed steps; per-step workspace ~1 GB >> L2",
            "parallelism": f"dp{world}: whole meshes per rank, no forward collective"}


def cpu_reference_run(steps: int, warmup: int, n_meshes: int = 1, n_vtx: int = N_VTX):
    """Reference CPU path (oracle port = op-for-op restatement of models/rignet.py pinned to the unmodified
    reference) on all host threads.  One step = forward of `n_meshes` 4096-vertex meshes."""
    from morig_b200 import synth
    from oracle import rignet_port
    import morig_b200
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = synth.ARCH_KWARGS[ARCH]
    model = getattr(morig_b200, ARCH)(**kw).eval()
    sd = synth.seeded_state_dict(model, 1)
    data = synth.make_batch(n_meshes, n_vtx, seed=0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            rignet_port.jointnet_motion_forward(sd, data, data.pred_flow, num_keyframes=kw["num_keyframes"],
                                                aggr_method=kw["aggr_method"])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    return dict(value=n_meshes * len(times) / total, unit=UNIT, cores=cores, kind="port",
                sample=f"{len(times)} timed forwards of {n_meshes} x {n_vtx}-vertex mesh after {warmup} warm-up, "
                       f"torch {torch.__version__} CPU fp32, {cores} threads",
                ms_per_step=1e3 * total / len(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step = the same batch our arm steps at N=1 (4 x 4096 vertices, ~3 s of CPU work): a bounded sample of the
    # N-GPU workload, which is N such batches
    r = cpu_reference_run(args.steps, args.warmup, n_meshes=args.meshes_per_gpu)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, args.meshes_per_gpu),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    import morig_b200
    from morig_b200 import _lib, engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    json_fd = 1
    if world > 1:
        # stdout must carry the ONE JSON line only: NCCL prints its version banner (and, with NCCL_DEBUG=INFO in the
        # environment, its whole log) to fd 1 from C, so fd 1 is pointed at stderr for the run and the line goes to
        # the saved descriptor
        # (NCCL_DEBUG is left as the caller set it: the driver audits the NCCL log for the rank count)
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()                                       # fail loudly if the extension is missing
    peaks = load_peaks()

    kw = synth.ARCH_KWARGS[ARCH]
    model = getattr(morig_b200, ARCH)(**kw).eval()
    model.load_state_dict(synth.seeded_state_dict(model, 1))
    model = model.to(dev)
    MESHES_PER_GPU = args.meshes_per_gpu
    host = synth.make_batch(MESHES_PER_GPU, N_VTX, seed=rank * MESHES_PER_GPU).pin_memory()
    resident = host.to(dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def fresh_views(b):
        # new tensor objects every step, like `data.to(device)` per iteration: graph prep is part of the step
        return synth.Batch(**{k: (v.view_as(v) if torch.is_tensor(v) else v) for k, v in b.__dict__.items()})

    counter = engine.LaunchCounter()
    prof = engine.KernelTimer()

    def step_resident():
        d = fresh_views(resident)
        return model(d, d.pred_flow)

    def step_e2e():
        d = host.to(dev, non_blocking=True)
        out = model(d, d.pred_flow)
        return [o.to("cpu", non_blocking=True) for o in out]

    def timed(fn, steps, warmup, profile=False):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            evs = []
            if profile:
                torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the timed steps only
            for _ in range(steps):
                flush_buf.fill_(1)                    # L2 flush between timed iterations
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                fn()
                e.record(stream)
                evs.append((s, e))
            barrier()
            if profile:
                torch.cuda.profiler.stop()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # timed region: after two sightings of the batch shape the module replays the whole forward (graph
    # preparation included) as one CUDA graph on static input buffers, so warm-up >= 3 covers the capture
    ms_total = timed(step_resident, args.steps, args.warmup, profile=os.environ.get("MORIG_BENCH_PROFILE") == "graph")
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e_serial = timed(step_e2e, args.steps, args.warmup)
    pipe = morig_b200.HostPipeline(model, depth=2)

    def timed_pipeline(steps, warmup):
        with torch.no_grad():
            for _ in range(max(warmup, len(pipe.slots) + 1)):      # warm up with the slots in flight, as in the timed loop
                if pipe.in_flight == len(pipe.slots):
                    pipe.result()
                pipe.submit(host, host.pred_flow)
            while pipe.in_flight:
                pipe.result()
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            for _ in range(steps):
                flush_buf.fill_(1)                    # L2 flush between forwards (compute stream)
                if pipe.in_flight == len(pipe.slots):
                    pipe.result()                     # host reads the oldest step's outputs
                pipe.submit(host, host.pred_flow)
            while pipe.in_flight:
                pipe.result()
            pipe.join(stream)
            e.record(stream)
            barrier()
        t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_e2e = timed_pipeline(args.steps, args.warmup)
    # instrumented pass (same steps, plain stream launches): per-kernel CUDA events + launch count.  Events cannot
    # be placed between the nodes of a replayed graph, so the per-kernel durations come from this pass.
    engine.set_hooks(counter, prof)
    timed(step_resident, 1, 1)
    counter.reset(); prof.reset()
    ms_instr = timed(step_resident, args.steps, 0, profile=os.environ.get("MORIG_BENCH_PROFILE") == "eager")
    launches = counter.count
    kstats = prof.summary()
    engine.set_hooks(None, None)

    # cached-graph variant (same Batch object re-submitted): informational
    def step_cached():
        return model(resident, resident.pred_flow)
    ms_cached = timed(step_cached, args.steps, args.warmup)

    # BASELINE.json configs[4] as stated (64 meshes over 8 GPUs = 8 meshes per GPU): an extra timed loop with that batch
    # whenever the job is multi-GPU; the headline stays on 4 meshes per GPU at every N so that the weak-scaling series
    # compares like with like
    cfg4 = None
    if world > 1 and MESHES_PER_GPU != 8:
        res8 = synth.make_batch(8, N_VTX, seed=1000 + rank * 8).to(dev)

        def step8():
            d = fresh_views(res8)
            return model(d, d.pred_flow)
        ms8 = timed(step8, args.steps, args.warmup)
        cfg4 = {"workload": f"{ARCH} forward, batch={8 * world} x {N_VTX}-vertex meshes sharded {8} per GPU over {world} GPUs "
                            "(BASELINE.json configs[4] at 8 GPUs)", "value": 8 * world * args.steps / (ms8 / 1e3), "unit": UNIT,
                "ms_per_step": ms8 / args.steps, "meshes_per_gpu": 8}
        del res8

    # training step of the same network and batch: forward + backward through this package's
    # kernels and ONE flat-buffer gradient all-reduce (NCCL over NVLink at N > 1, dp.GradAllReduce) -- informational,
    # outside the headline metric
    train = None
    if args.train_steps > 0:
        from morig_b200 import dp
        tmodel = getattr(morig_b200, ARCH)(**kw)
        tmodel.load_state_dict(synth.seeded_state_dict(tmodel, 1))
        tmodel = tmodel.to(dev).train()
        ar = dp.GradAllReduce(tmodel)
        target = torch.zeros(resident.pos.shape[0], 3, device=dev)

        def train_step():
            ar.zero_grad()
            _, _, pred = tmodel(resident, resident.pred_flow)
            loss = (torch.tanh(pred) - target).pow(2).mean()       # stand-in for training/train_rig.py:168-183
            loss.backward()
            return ar.finish()

        nbytes = train_step()
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record(stream)
        for _ in range(args.train_steps):
            train_step()
        e_ev.record(stream)
        barrier()
        t = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_train = float(t.item()) / args.train_steps
        train = {"ms_per_step": ms_train, "meshes_per_s": MESHES_PER_GPU * world / (ms_train / 1e3),
                 "steps": args.train_steps, "allreduce_bytes_per_step": nbytes, "buckets": len(ar.buckets),
                 "what": "forward + backward (train-mode BatchNorm; forward / input-gradient GEMMs on the tcgen05 split-fp16 engine, weight gradients 3xTF32 tcgen05) + flat-buffer gradient all-reduce, "
                         "max over ranks; BatchNorm statistics per replica",
                 "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
        ar.close()
        del tmodel, ar
        torch.cuda.empty_cache()

    if rank == 0:
        meshes = MESHES_PER_GPU * world * args.steps
        value = meshes / (ms_total / 1e3)
        h2d = sum(v.numel() * v.element_size() for v in host.__dict__.values() if torch.is_tensor(v))
        with torch.no_grad():
            outs = step_resident()
        d2h = sum(o.numel() * o.element_size() for o in outs)
        alg_flops_step = ALG_FLOP_PER_VERTEX * N_VTX * MESHES_PER_GPU
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
        if kstats and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(kstats[0]["kernel"])      # ncu dram bytes per launch of the dominant kernel
        roof = engine.roofline_report(kstats, peaks, ms_total / args.steps, alg_flops_step, traffic)
        # the memory-leaning member of the fused EdgeConv family (C_x = 3 layers, north_star's HBM-roofline kernel): reported
        # against the measured HBM peak next to the tensor-bound dominant kernel; see DESIGN.md 5.2 for why it stays low
        narrow = next((k for k in kstats if k["kernel"].startswith("edgeconv H=32") and "frames=5" in k["kernel"]), None)
        roof_hbm = None
        if narrow is not None:
            tr = None
            if os.path.exists(tpath):
                with open(tpath) as f:
                    tr = json.load(f).get(narrow["kernel"])
            roof_hbm = {"bound": "hbm", "kernel": narrow["kernel"], "achieved": narrow["gbs"], "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": narrow["gbs"] / peaks["hbm_gbs"], "traffic": tr,
                        "avg_launch_ms": narrow["avg_ms"], "alg_mb_per_launch": narrow["alg_mb_per_launch"],
                        # ncu (profiles/r02_ncu_dominant_kernels.txt): L2 read sectors from L1/TEX per launch
                        "l2_read_mb_per_launch_ncu": 151.7,
                        "l2_gbs_ncu": round(151.7e-3 / (narrow["avg_ms"] * 1e-3), 1) if narrow["avg_ms"] else None,
                        "l2_cap_gbs": 12400.0,
                        "note": "86 FLOP/B and 1.39 M random 128-byte gathers per launch: neither HBM (6 %) nor L2 bandwidth "
                                "(152 MB of L2 reads per launch = ~1.7 TB/s = 14 % of the ~12.4 TB/s L2->SM cap) binds; ncu: issue "
                                "slots 54 % busy at 16 resident warps per SM -- instruction issue and gather / shuffle latency of "
                                "the warp-level MMA kernel (more occupancy through a register cap was measured: slower)"}
        cpu = cpu_reference_run(steps=3, warmup=1, n_meshes=1) if world == 1 else None
        # BASELINE.json configs[0] (1 x 1024 vertices, the reference's own CPU-runnable case): both sides, informational
        cfg0 = None
        if world == 1:
            small = synth.make_batch(1, 1024, seed=0).to(dev)
            ms0 = timed(lambda: model(small, small.pred_flow), args.steps, max(args.warmup, 3))
            c0 = cpu_reference_run(steps=5, warmup=1, n_meshes=1, n_vtx=1024)
            cfg0 = {"workload": "jointnet_motion forward, 1 x 1024-vertex mesh (BASELINE.json configs[0])",
                    "gpu_ms_per_forward": ms0 / args.steps, "gpu_meshes_per_s": args.steps / (ms0 / 1e3),
                    "cpu_ms_per_forward": c0["ms_per_step"], "cpu_meshes_per_s": c0["value"], "cpu_cores": c0["cores"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world, MESHES_PER_GPU),
                "clocks": clocks,
                "e2e": {"value": meshes / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                        "api": "morig_b200.HostPipeline(model, depth=2): pinned host batch in, pinned host outputs "
                               "out, every step; copies of neighbouring steps overlap the forward",
                        "serial": {"value": meshes / (ms_e2e_serial / 1e3), "ms_per_step": ms_e2e_serial / args.steps,
                                   "api": "model(data.to(dev), flow) then .to('cpu'), one step after the other"}},
                "gpu_launches": launches,
                "launch_mode": "timed steps replay one CUDA graph per step (same kernels, graph preparation included); "
                               "gpu_launches / kernels / roofline come from an instrumented pass of the same "
                               f"{args.steps} steps with plain stream launches ({ms_instr / args.steps:.3f} ms/step)",
                "cached_graph": {"value": meshes / (ms_cached / 1e3), "unit": UNIT,
                                 "note": "same Batch object re-submitted: CSR cache hit, informational"},
                "roofline": roof,
                "roofline_narrow_edgeconv": roof_hbm,
                "model_tflops_algorithmic": alg_flops_step * world / (ms_total / args.steps / 1e3) / 1e12,
                "kernels": kstats}
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if cfg0 is not None:
            line["config0_1x1024"] = cfg0
        if train is not None:
            line["train_step"] = train
        if cfg4 is not None:
            line["configs4_8_per_gpu"] = cfg4
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-steps", type=int, default=3,
                    help="extra: timed training steps (forward + backward + gradient all-reduce) reported under "
                         "`train_step`; 0 disables")
    ap.add_argument("--meshes-per-gpu", type=int, default=MESHES_PER_GPU,
                    help="meshes in one step's batch per GPU (4 = BASELINE.json configs[1]; 8 at --gpus 8 = configs[4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
