#!/bin/bash
# A/B of experiment builds (morig_b200/libmorig_b200_exp*.so, built with MORIG_NVCC_FLAGS=... MORIG_LIB=...) against the shipped
# library: short bench each, headline + the kernels matching $1 (grep -E pattern)
set -u
mkdir -p gpurun_out
pat="${1:-^dense|^edgeconv H=(128|256)}"
for lib in base morig_b200/libmorig_b200_exp*.so; do
  if [ "$lib" != base ]; then export MORIG_LIB=$PWD/$lib; fi
  v=$(basename $lib .so)
  python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  echo "== $v"; python scripts/show_bench.py gpurun_out/bench_ab_$v.json 45 | grep -E "^\{'value|$pat"
done
