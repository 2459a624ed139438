#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -p no:cacheprovider > gpurun_out/t_conv.log 2>&1; echo "conv rc $?"; tail -1 gpurun_out/t_conv.log
timeout 200 python scripts/profile_layers.py --precision f16f8 > gpurun_out/layers_f16f8_s.txt 2> gpurun_out/lay.err; head -1 gpurun_out/layers_f16f8_s.txt
timeout 200 python scripts/profile_layers.py --precision bf16x3 > gpurun_out/layers_bf16x3_s.txt 2>> gpurun_out/lay.err; head -1 gpurun_out/layers_bf16x3_s.txt
MPN_RES_MMA=0 timeout 200 python scripts/profile_layers.py --precision bf16x3 > gpurun_out/layers_bf16x3_s_nores.txt 2>> gpurun_out/lay.err; head -1 gpurun_out/layers_bf16x3_s_nores.txt
timeout 300 python scripts/exp_dual_stream.py --precision f16f8 > gpurun_out/dual_f16f8.txt 2> gpurun_out/dual.err; cat gpurun_out/dual_f16f8.txt | head -4
timeout 300 python scripts/exp_dual_stream.py --precision bf16x3 --splits 1,2 > gpurun_out/dual_bf16x3.txt 2>> gpurun_out/dual.err; cat gpurun_out/dual_bf16x3.txt | head -3
timeout 400 python scripts/parity_margin.py > gpurun_out/parity_margin.txt 2> gpurun_out/parity.err; tail -1 gpurun_out/parity.err; grep -c precision gpurun_out/parity_margin.txt
