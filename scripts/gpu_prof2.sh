set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_kernel" -o /tmp/prof_k2 python scripts/prof_kernels.py e256 e128 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-200
cp /tmp/prof_k2.ncu-rep gpurun_out/prof_k2.ncu-rep
ls -la gpurun_out/prof_k2.ncu-rep
