#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -3 gpurun_out/t_all_gpu.log | cut -c1-300
grep -E "^FAILED|^ERROR|Error" gpurun_out/t_all_gpu.log | head -10
timeout 500 python bench.py --no-cpu-baseline --no-fast > gpurun_out/bench_y1.json 2> gpurun_out/bench_y1.err; echo "bench rc $?"
python - <<'PY'
import json
for f in ("bench_y1",):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("e2e_u8_input") or {}).get("value"),
      (d.get("roofline") or {}).get("frac"), d.get("clocks"))
PY
