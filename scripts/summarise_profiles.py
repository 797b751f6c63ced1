"""Turns the raw ncu exports of scripts/gpu_profile.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

# ---- launch list of the graph-replayed steps ----
rows = list(csv.reader(open(os.path.join(OUT, "launches_graph.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    v, unit = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
    us = v / 1000.0 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1000.0)
    a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += us
    tot += us
lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 (the 2 timed, graph-replayed steps)",
         f"# {len(data)} kernel launches in 2 steps, sum of durations {tot / 2:.1f} us per step (cold-cache, serialised: shares matter, not absolutes)",
         f"# {'launches':>8s} {'total_us':>10s} {'share':>6s}  kernel"]
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"  {n:8d} {us:10.1f} {100 * us / tot:5.1f}%  {name[:150]}")
open(os.path.join(PROF, "r02_launches_graph_step.txt"), "w").write("\n".join(lines) + "\n")

# ---- full-set summary of the dominant kernels ----
rows = list(csv.reader(open(os.path.join(OUT, "prof_kernels_raw.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from L1/TEX)"),
        ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "L1 global-load bytes")]
labels = ["edgeconv H=256 E=278528 N=16384 frames=5", "edgeconv H=128 E=278528 N=16384 frames=5", "dense M=81920 N=1024 K=840",
          "edgeconv H=32 E=278528 N=16384 frames=5", "edgeconv H=16 E=278528 N=16384 frames=1"]
out = ["# ncu --set full --clock-control none --import-source on, one launch of each dominant kernel shape (scripts/prof_kernels.py),",
       "# fp16-split operand kind, 4 x 4096-vertex jointnet_motion shapes.  Report: gpurun_out/prof_kernels.ncu-rep (not committed, 15 MB)", ""]
mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
traffic = {}
for k, r in enumerate(data):
    out.append(f"## {labels[k] if k < len(labels) else k}: {r[ix['Kernel Name']][:110]}")
    for m, lab in want:
        if m in ix:
            out.append(f"  {lab:28s} {r[ix[m]]} {units[ix[m]]}")
    rd = float(r[ix["dram__bytes_read.sum"]]) * mul[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]]) * mul[units[ix["dram__bytes_write.sum"]]]
    if k < len(labels):
        traffic[labels[k]] = int(rd + wr)
    out.append("")
# the narrow kernel again with the per-lane LDG gather (MORIG_NARROW_TMA=0) for the A/B of the TMA staging
p2 = os.path.join(OUT, "prof_narrow_ldg_raw.csv")
if os.path.exists(p2):
    rows2 = list(csv.reader(open(p2)))
    hdr2, units2, data2 = rows2[0], rows2[1], rows2[2:]
    ix2 = {h: i for i, h in enumerate(hdr2)}
    for r in data2[:1]:
        out.append(f"## edgeconv H=32 E=278528 N=16384 frames=5, per-lane LDG.128 gather (MORIG_NARROW_TMA=0): {r[ix2['Kernel Name']][:90]}")
        for m, lab in want:
            if m in ix2:
                out.append(f"  {lab:28s} {r[ix2[m]]} {units2[ix2[m]]}")
        out.append("")
open(os.path.join(PROF, "r02_ncu_dominant_kernels.txt"), "w").write("\n".join(out))
json.dump(traffic, open(os.path.join(PROF, "r02_dram_traffic.json"), "w"), indent=1)
print("\n".join(lines[:14]))
print(traffic)
