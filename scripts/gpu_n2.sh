#!/bin/bash
# Two-GPU pass:  gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_n2.sh'
# the DataParallel replica test on two devices, then the default bench line under torchrun (inference shards without a collective;
# train_step with the NCCL gradient allreduce at 2 ranks; full_pipeline at 2 ranks)
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider -k "dataparallel or train_step or shard" > gpurun_out/t_gpu_n2.log 2>&1; echo "exit n2 tests: $?"; tail -3 gpurun_out/t_gpu_n2.log | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc $?"; tail -2 gpurun_out/bench_n2.err | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print("bench n2", {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", (d.get("e2e") or {}).get("value"))
    for k in ("train_step", "full_pipeline"):
        v = d.get(k) or {}
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "n_gpus", "allreduce_alone_ms", "error") if v.get(kk) is not None}, (v.get("config") or {}).get("allreduce_overlapped"))
except Exception as e:
    print("unreadable", e)
PY
