mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -q -m gpu --timeout=600 -p no:cacheprovider -s > gpurun_out/t_train.log 2>&1; echo "exit train tests: $?"
grep -E "passed|failed|graph vs eager|checked|FAILED" gpurun_out/t_train.log
python scripts/profile_train.py 16 bf16x3 > gpurun_out/profile_train_b16.txt 2>&1; sed -n 3,10p gpurun_out/profile_train_b16.txt
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; cat gpurun_out/bench_train_n1.json | cut -c1-330
