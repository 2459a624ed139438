mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
   python bench.py --mode train --steps 1 --warmup 3 --batch 8 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --mode train --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc $?"
cat gpurun_out/bench_train_n2.json; tail -3 gpurun_out/bench_train_n2.err
