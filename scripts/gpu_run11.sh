mkdir -p gpurun_out
timeout 600 python scripts/debug_train2.py > gpurun_out/debug_train2.log 2>&1; echo "exit: $?"
tail -40 gpurun_out/debug_train2.log | cut -c1-200
MPN_SPLIT_BN256=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout=300 -p no:cacheprovider -k "bf16x3" > gpurun_out/t_conv_bn256.log 2>&1; echo "exit conv bn256: $?"; tail -5 gpurun_out/t_conv_bn256.log
