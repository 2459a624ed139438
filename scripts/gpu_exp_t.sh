#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -p no:cacheprovider > gpurun_out/t_conv.log 2>&1; echo "conv rc $?"; tail -1 gpurun_out/t_conv.log
grep -E "^FAILED|^ERROR|AssertionError: err|mbarrier|MpnError|CUDA error" gpurun_out/t_conv.log | head -10
MPN_RES_MMA=0 timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -p no:cacheprovider -k "bf16x3" > gpurun_out/t_conv_nores.log 2>&1; echo "conv (RES_MMA=0) rc $?"; tail -1 gpurun_out/t_conv_nores.log
grep -E "^FAILED|^ERROR|AssertionError: err|mbarrier|MpnError|CUDA error" gpurun_out/t_conv_nores.log | head -10
timeout 200 python scripts/profile_layers.py --precision f16f8 > gpurun_out/layers_f16f8_t.txt 2> gpurun_out/lay.err; head -1 gpurun_out/layers_f16f8_t.txt; grep " res" gpurun_out/layers_f16f8_t.txt
MPN_RES_TMA=0 timeout 200 python scripts/profile_layers.py --precision f16f8 > gpurun_out/layers_f16f8_t_lsu.txt 2> gpurun_out/lay.err; head -1 gpurun_out/layers_f16f8_t_lsu.txt; grep " res" gpurun_out/layers_f16f8_t_lsu.txt
timeout 200 python scripts/profile_layers.py --precision bf16x3 > gpurun_out/layers_bf16x3_t.txt 2>> gpurun_out/lay.err; head -1 gpurun_out/layers_bf16x3_t.txt; grep " res" gpurun_out/layers_bf16x3_t.txt
MPN_RES_MMA=0 timeout 200 python scripts/profile_layers.py --precision bf16x3 > gpurun_out/layers_bf16x3_t_nores.txt 2>> gpurun_out/lay.err; head -1 gpurun_out/layers_bf16x3_t_nores.txt; grep " res" gpurun_out/layers_bf16x3_t_nores.txt
timeout 300 python -m pytest tests/test_gpu_network.py -x -q -p no:cacheprovider > gpurun_out/t_net.log 2>&1; echo "net rc $?"; tail -1 gpurun_out/t_net.log
