#!/bin/bash
# fp16-split engine bring-up: targeted kernel tests (under timeout: a pipeline bug traps instead of hanging), whole
# suite, model error, microbench of the three engines, bench with both operand kinds
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "tensor_core or fp16 or test_dense_fwd" 2>&1 | tail -40 | tee gpurun_out/pytest_tc.txt
timeout 900 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 python scripts/model_err.py 2>&1 | tail -4 | tee gpurun_out/model_err.txt
MORIG_TC_KIND=tf32 timeout 300 python scripts/model_err.py 2>&1 | tail -4 | tee gpurun_out/model_err_tf32.txt
timeout 400 python scripts/tc_microbench.py 2>&1 | tail -20 | tee gpurun_out/tc_microbench.txt
MORIG_TC_KIND=tf32 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_tf32.json', 'gpurun_out/bench.json'):
    try:
        d=json.load(open(f)); ks=d.pop('kernels')
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in d})
    for k in ks[:16]: print('  ', k['kernel'], k['avg_ms'], k['tflops'])
PY
tail -3 gpurun_out/bench.err
