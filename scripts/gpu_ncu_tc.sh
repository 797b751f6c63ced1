#!/bin/bash
set -u
mkdir -p gpurun_out
for case in "t.dense_case(81920, 256, 1024)" "t.edge_case(128, 5)" "t.edge_case(256, 5)"; do
name=$(echo "$case" | tr -c 'a-z0-9' '_' | cut -c1-24)
cat > /tmp/one.py <<PY
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
$case
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_$name python /tmp/one.py > gpurun_out/ncu_$name.log 2>&1
tail -1 gpurun_out/ncu_$name.log
done
ls -la gpurun_out/*.ncu-rep
