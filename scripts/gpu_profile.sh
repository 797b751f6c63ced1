#!/bin/bash
# Evidence run: launch list of the timed (graph-replayed) steps + full ncu capture of the tensor-core / edge kernels
# of one instrumented step.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
MORIG_BENCH_PROFILE=graph timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/launches_graph.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
MORIG_BENCH_PROFILE=eager timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_kernel|edge_mma" -c ${1:-70} -o gpurun_out/prof_bench_tc python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/
