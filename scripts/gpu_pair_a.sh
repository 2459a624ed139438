#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -p no:cacheprovider -k "tcgen05" > gpurun_out/t_pair_conv.log 2>&1; echo "conv pair rc $?"
grep -E "passed|failed" gpurun_out/t_pair_conv.log | tail -2; grep -E "^FAILED|^ERROR|AssertionError: err|mbarrier|MpnError|CUDA error" gpurun_out/t_pair_conv.log | head -20
