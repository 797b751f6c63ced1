#!/bin/bash
# Iteration round trip: GPU parity tests, role timelines of the fused EdgeConv kernels, short bench (no training step)
set -u
mkdir -p gpurun_out
export MORIG_BUILD_INCREMENTAL=1
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -k "${1:-not training}" 2>&1 | tail -15 > gpurun_out/pytest_iter.txt
tail -6 gpurun_out/pytest_iter.txt
for c in edge128 edge256; do
MORIG_LIB=$PWD/morig_b200/libmorig_b200_trace.so timeout 120 python scripts/tc_trace.py $c f16 5 > gpurun_out/trace_$c.txt 2>&1
grep -E "^edge|^\{" gpurun_out/trace_$c.txt | cut -c1-400
done
timeout 600 python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python scripts/show_bench.py gpurun_out/bench_iter.json 2>/dev/null | head -30
