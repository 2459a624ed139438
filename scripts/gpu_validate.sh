#!/bin/bash
# One-shot validation on a B200 box (what the round-end driver runs, plus the profiling helpers):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -3 gpurun_out/t_all_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; echo "train rc $?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref", "bench_train_n1"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("e2e_u8_input") or {}).get("value"),
          (d.get("roofline") or {}).get("frac"))
PY
