mkdir -p gpurun_out
timeout 900 python scripts/debug_train.py > gpurun_out/debug_train.log 2>&1; echo "exit: $?"
tail -150 gpurun_out/debug_train.log | cut -c1-200
