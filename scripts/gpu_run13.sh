mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -q -m gpu --timeout=600 -p no:cacheprovider -s > gpurun_out/t_train_step.log 2>&1; echo "exit train_step: $?"
grep -E "checked|max-norm|passed|failed|Error|assert" gpurun_out/t_train_step.log | head -20
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; echo "bench train rc $?"
cat gpurun_out/bench_train_n1.json; tail -5 gpurun_out/bench_train_n1.err
