mkdir -p gpurun_out
timeout 600 python scripts/debug_train3.py > gpurun_out/debug_train3.log 2>&1; tail -5 gpurun_out/debug_train3.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -q -m gpu --timeout=600 -p no:cacheprovider -s > gpurun_out/t_train.log 2>&1; echo "exit train tests: $?"
grep -E "passed|failed|graph vs eager|checked|FAILED" gpurun_out/t_train.log
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; cat gpurun_out/bench_train_n1.json | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 30 -c 3 -o gpurun_out/prof_wgrad_r01g python scripts/profile_train.py 8 bf16x3 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc $?"
