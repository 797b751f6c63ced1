#!/bin/bash
# TC engine bring-up: targeted tests first (under timeout so a pipeline bug cannot hang the box), then microbench + bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -x -k "tensor_core" 2>&1 | tail -8 | tee gpurun_out/pytest_tc.txt
timeout 600 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.txt
timeout 300 python scripts/model_err.py 2>&1 | tail -4 | tee gpurun_out/model_err.txt
timeout 300 python scripts/tc_microbench.py 2>&1 | tail -20 | tee gpurun_out/tc_microbench.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json')); ks=d.pop('kernels')
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cached_graph','cpu_baseline') if k in d})
for k in ks[:14]: print(k['kernel'], k['avg_ms'], k['tflops'])
PY
tail -3 gpurun_out/bench.err
