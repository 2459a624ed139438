#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 8 -f -o gpurun_out/r01s_conv_f16f8 \
  python scripts/bench_conv_shapes.py --precision f16f8 --once --only "l3.conv3,l3.conv1,l3.conv2,l1.conv3" > gpurun_out/ncu_u.log 2>&1
echo "ncu rc $?"; tail -8 gpurun_out/ncu_u.log
ls -la gpurun_out/*.ncu-rep
