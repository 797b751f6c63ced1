#!/bin/bash
# ncu --set full (source attached) of one launch of the fused EdgeConv shapes given as arguments (default e128 e256)
set -u
mkdir -p gpurun_out
which="${*:-e128 e256}"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_kernel|edge_mma" -f -o gpurun_out/prof_edge python scripts/prof_kernels.py $which > gpurun_out/ncu_edge.log 2>&1
tail -2 gpurun_out/ncu_edge.log | cut -c1-200
ls -la gpurun_out/prof_edge.ncu-rep
