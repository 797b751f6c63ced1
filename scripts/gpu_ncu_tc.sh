#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/one.py <<PY
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
t.edge_case(256, 5)
PY
MORIG_NO_2CTA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_e256 python /tmp/one.py > gpurun_out/ncu_e256.log 2>&1
tail -1 gpurun_out/ncu_e256.log
