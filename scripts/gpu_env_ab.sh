#!/bin/bash
# A/B of an environment switch of the library ($1, e.g. MORIG_EXP_A3=1): short bench without / with it; $2 = grep -E pattern
set -u
mkdir -p gpurun_out
pat="${2:-^edgeconv H=(128|256)}"
python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_env_base.json 2> gpurun_out/bench_env_base.err
echo "== base"; python scripts/show_bench.py gpurun_out/bench_env_base.json 45 | grep -E "^\{'value|$pat"
env "$1" python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_env_exp.json 2> gpurun_out/bench_env_exp.err
echo "== $1"; python scripts/show_bench.py gpurun_out/bench_env_exp.json 45 | grep -E "^\{'value|$pat"
