#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -x -q -p no:cacheprovider > gpurun_out/t_net.log 2>&1; echo "net rc $?"; tail -1 gpurun_out/t_net.log
grep -E "^FAILED|^ERROR|Error|assert" gpurun_out/t_net.log | head -10
timeout 500 python bench.py --no-cpu-baseline --no-fast > gpurun_out/bench_y1.json 2> gpurun_out/bench_y1.err; echo "bench rc $?"
python - <<'PY'
import json
for f in ("bench_y1",):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("e2e_u8_input") or {}).get("value"),
      (d.get("roofline") or {}).get("frac"), d.get("clocks"))
PY
