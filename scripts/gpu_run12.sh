mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -q -m gpu --timeout=600 -p no:cacheprovider -s > gpurun_out/t_train_step.log 2>&1; echo "exit train_step: $?"
grep -E "checked|max-norm|passed|failed|Error|assert" gpurun_out/t_train_step.log | head -20
