mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -q -m gpu --timeout=600 -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "exit train tests: $?"
tail -5 gpurun_out/t_train.log; grep -E "Error|error" gpurun_out/t_train.log | head -5
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; echo "bench train rc $?"
cat gpurun_out/bench_train_n1.json; tail -3 gpurun_out/bench_train_n1.err
python bench.py --mode train --steps 10 --warmup 3 --precision bf16 > gpurun_out/bench_train_n1_bf16.json 2> gpurun_out/bench_train_n1_bf16.err; echo "bench train bf16 rc $?"
cat gpurun_out/bench_train_n1_bf16.json
