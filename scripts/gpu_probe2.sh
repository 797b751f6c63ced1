set -u
mkdir -p gpurun_out
MORIG_LIB=$PWD/morig_b200/libmorig_b200_trace.so timeout 120 python scripts/tc_trace.py dense f16 1 > gpurun_out/trace_dense.txt 2>&1
grep -E "^dense|^\{" gpurun_out/trace_dense.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fill_cut" 2>&1 | tail -2
