#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prn_assign.py tests/test_gpu_pipeline.py -x -q -p no:cacheprovider > gpurun_out/t_prn.log 2>&1; echo "prn tests rc $?"; tail -1 gpurun_out/t_prn.log
timeout 300 python bench.py --mode full --steps 5 --warmup 3 > gpurun_out/bench_full_z.json 2> gpurun_out/bench_full_z.err; echo "full rc $?"; tail -2 gpurun_out/bench_full_z.err; cat gpurun_out/bench_full_z.json | cut -c1-900
timeout 300 python scripts/profile_full.py > gpurun_out/profile_full.txt 2>&1; head -8 gpurun_out/profile_full.txt
