#!/bin/bash
# bench at N GPUs of one box (driver-style launch) + the reference arm
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json | cut -c1-1500; tail -3 gpurun_out/bench_n$N.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-900
