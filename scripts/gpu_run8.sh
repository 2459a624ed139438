mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -q -m gpu --timeout=300 -p no:cacheprovider -x > gpurun_out/t_train_ops.log 2>&1; echo "exit train_ops(-x): $?"
tail -40 gpurun_out/t_train_ops.log
timeout 900 python -m pytest tests/test_gpu_train_ops.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_train_ops_all.log 2>&1; echo "exit train_ops(all): $?"
grep -E "passed|failed|FAILED|Error" gpurun_out/t_train_ops_all.log | head -40
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout=300 -p no:cacheprovider -k "prn or golden" > gpurun_out/t_network.log 2>&1; echo "exit network: $?"; tail -4 gpurun_out/t_network.log
