#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ncu_conv_step.py --precision f16f8 --steps 2 --names-out gpurun_out/conv_launch_names.txt > gpurun_out/names.log 2>&1; echo "names rc $?"; wc -l gpurun_out/conv_launch_names.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 3 -f -o gpurun_out/${NCU_TAG:-r02}_conv_f16f8 \
  python scripts/bench_conv_shapes.py --precision f16f8 --once --only "${NCU_ONLY:-l3.conv3,l3.conv1,l3.conv2}" > gpurun_out/ncu_u.log 2>&1
echo "ncu rc $?"; tail -5 gpurun_out/ncu_u.log
ls -la gpurun_out/*.ncu-rep
