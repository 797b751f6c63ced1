#!/bin/bash
# One GPU-box round trip: parity tests, smoke, short bench, ncu launch list.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt | tail -15
python __graft_entry__.py smoke 2>&1 | tail -6 | tee gpurun_out/smoke.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi
