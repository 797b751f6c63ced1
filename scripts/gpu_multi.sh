#!/bin/bash
# bench at N GPUs of one box (driver-style launch); NCCL log excerpt; ragged strong-scaling line
N=${1:-2}
mkdir -p gpurun_out
export MORIG_BUILD_INCREMENTAL=1
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1200 gpurun_out/bench_n$N.json | cut -c1-1200; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
grep -E "NCCL INFO (Connected|Channel|comm|ncclCommInitRank|Using network|NVLS)" gpurun_out/bench_n$N.err | head -12 > gpurun_out/nccl_excerpt_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/scale_ragged.py > gpurun_out/ragged_n$N.json 2> gpurun_out/ragged_n$N.err
tail -c 400 gpurun_out/ragged_n$N.json
