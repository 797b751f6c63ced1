#!/bin/bash
# A/B of an experiment build (MORIG_LIB=morig_b200/libmorig_b200_exp.so) against the shipped library: short bench each
set -u
mkdir -p gpurun_out
for v in base exp; do
  if [ $v = exp ]; then export MORIG_LIB=$PWD/morig_b200/libmorig_b200_exp.so; fi
  python bench.py --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  echo "== $v"; python scripts/show_bench.py gpurun_out/bench_ab_$v.json 12 | grep -E "^\{'value|^dense|^edgeconv H=(128|256)"
done
