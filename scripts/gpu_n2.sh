#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-fast > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc $?"; tail -1 gpurun_out/bench_n2.json | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --mode train --steps 8 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc $?"; tail -1 gpurun_out/bench_train_n2.json | cut -c1-400
