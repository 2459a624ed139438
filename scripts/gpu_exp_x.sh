#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -3 gpurun_out/t_all_gpu.log | cut -c1-300
grep -E "^FAILED|^ERROR|Error" gpurun_out/t_all_gpu.log | head -10
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_x.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "dtype")}, (d.get("e2e") or {}).get("value"), (d.get("e2e_u8_input") or {}).get("value"),
      (d.get("roofline") or {}).get("frac"), (d.get("alt_parity_mode") or {}).get("value"), (d.get("fast_mode") or {}).get("value"), d.get("clocks"))
PY
