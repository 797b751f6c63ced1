#!/bin/bash
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import tc_microbench as t
t.dense_case(81920, 832, 1024)
t.dense_case(81920, 256, 1024)
t.edge_case(256, 5)
t.edge_case(128, 5)
PY
for cfg in "MORIG_NO_2CTA=1 MORIG_TC_DBG=0" "MORIG_NO_2CTA=0 MORIG_TC_DBG=0" "MORIG_NO_2CTA=1 MORIG_TC_DBG=4" "MORIG_NO_2CTA=1 MORIG_TC_DBG=7"; do echo "== $cfg"; env $cfg timeout 120 python /tmp/one.py 2>&1 | grep -o "^[a-z]* [^:]*: \|tc [0-9.]* ms [0-9.]* TF/s\|rror.*" | paste - - ; done
