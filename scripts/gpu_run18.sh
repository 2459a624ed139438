mkdir -p gpurun_out
python scripts/profile_train.py 16 bf16x3 > gpurun_out/profile_train_b16.txt 2>&1; tail -27 gpurun_out/profile_train_b16.txt
python scripts/profile_train.py 16 bf16 > gpurun_out/profile_train_b16_bf16.txt 2>&1; tail -14 gpurun_out/profile_train_b16_bf16.txt
