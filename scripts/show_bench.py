"""Pretty-print a bench.py JSON line: headline numbers and the per-kernel table."""
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"))
ks = d.pop("kernels", [])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d})
print("e2e", d.get("e2e"))
print("roofline", d.get("roofline"))
print("cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
tot = sum(k["total_ms"] for k in ks)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k in ks[:top]:
    print(f"{k['kernel']:52s} n={k['launches']:4d} avg={k['avg_ms']:.4f} share={k['total_ms'] / tot * 100:5.1f}% "
          f"tf={k['tflops']:.1f} gbs={k['gbs']:.0f}")
print("sum of kernel ms per step:", tot / max(d.get("steps", 1), 1))
