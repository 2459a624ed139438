#!/bin/bash
# pair-kernel validation: full GPU suite, smoke, bench in both parity modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -3 gpurun_out/t_all_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
timeout 300 python bench.py --precision f16f8 --no-cpu-baseline --no-fast --steps 10 > gpurun_out/bench_f8.json 2> gpurun_out/bench_f8.err; echo "bench f8 rc $?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_f8"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
