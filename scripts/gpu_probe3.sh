set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -x --tb=short -k "wgrad" 2>&1 | tail -15
python scripts/wgrad_bench.py 2>&1 | tail -2
