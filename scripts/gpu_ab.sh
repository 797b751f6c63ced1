#!/bin/bash
# A/B: cta_group::2 on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line -x 2>&1 | tail -5
for v in 0 1; do echo "== MORIG_NO_2CTA=$v"; MORIG_NO_2CTA=$v timeout 300 python scripts/tc_microbench.py 2>&1 | grep -v "K=96 \|M=16384" | grep -o "^[a-z]* [^:]*: \|tc [0-9.]* ms [0-9.]* TF/s" | paste - -
MORIG_NO_2CTA=$v timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('bench value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
