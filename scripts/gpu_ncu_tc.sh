#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
t.dense_case(81920, 832, 1024)
t.edge_case(256, 5)
t.edge_case(128, 5)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_tc_dense python /tmp/one.py > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
cat > /tmp/two.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
t.edge_case(256, 5)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_tc_edge256 python /tmp/two.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
cat > /tmp/three.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
t.edge_case(128, 5)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -o gpurun_out/prof_tc_edge128 python /tmp/three.py > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
