mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
for g in detect conv network; do
  timeout 600 python -m pytest tests/test_gpu_$g.py -q -m gpu -x --timeout=300 -p no:cacheprovider > gpurun_out/t_$g.log 2>&1
  echo "exit $g: $?" >> gpurun_out/summary.txt
done
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_conv_all.log 2>&1
echo "exit conv_all: $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_detect.log; tail -30 gpurun_out/t_conv.log
