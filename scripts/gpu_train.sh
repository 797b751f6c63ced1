set -u
mkdir -p gpurun_out
export MORIG_BUILD_INCREMENTAL=1
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -k "training or boundary" 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1])
print(d['value'], d['train_step'])
P
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python scripts/train_profile.py > gpurun_out/train_profile.log 2>&1
python scripts/sum_launches.py gpurun_out/train_launches.csv | head -24 | cut -c1-130 | tee gpurun_out/train_launches_summary.txt
