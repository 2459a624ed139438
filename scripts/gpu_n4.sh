mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4 rc $?"; tail -2 gpurun_out/bench_n4.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n4.json").read().strip().splitlines()[-1])
print("bench n4", {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", (d.get("e2e") or {}).get("value"))
for k in ("train_step", "full_pipeline"):
    v = d.get(k) or {}
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "n_gpus", "allreduce_alone_ms", "error") if v.get(kk) is not None}, (v.get("config") or {}).get("allreduce_overlapped"))
PY
