#!/bin/bash
set -u
mkdir -p gpurun_out
export MORIG_BUILD_INCREMENTAL=1
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -k "not training" 2>&1 | tail -4
MORIG_LIB=$PWD/morig_b200/libmorig_b200_trace.so timeout 120 python scripts/tc_trace.py dense f16 1 > gpurun_out/trace_dense.txt 2>&1
grep -E "^dense|^\{" gpurun_out/trace_dense.txt | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python scripts/show_bench.py gpurun_out/bench_iter.json 14 | grep -E "^\{'value|^dense|^edgeconv H=(128|256)"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
print('config0', d.get('config0_1x1024'))
P
