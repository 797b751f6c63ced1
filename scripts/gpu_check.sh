#!/bin/bash
# One GPU-box round trip: parity tests, smoke, short bench (both arms).  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
export MORIG_BUILD_INCREMENTAL=1          # the in-tree .so travels with the snapshot
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 > gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -6 | tee gpurun_out/smoke.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" = "ref" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  tail -c 600 gpurun_out/bench_ref.json
fi
