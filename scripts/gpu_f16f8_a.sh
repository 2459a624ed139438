#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -p no:cacheprovider -k "f16f8 or stem" > gpurun_out/t_f8_conv.log 2>&1; echo "conv f16f8 rc $?"
grep -E "passed|failed" gpurun_out/t_f8_conv.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/t_f8_conv.log | head -40
grep -E "AssertionError: err|mbarrier|MpnError" gpurun_out/t_f8_conv.log | head -20
timeout 600 python -m pytest tests/test_gpu_network.py -q -p no:cacheprovider -k "f16f8" > gpurun_out/t_f8_net.log 2>&1; echo "net f16f8 rc $?"
grep -E "passed|failed" gpurun_out/t_f8_net.log | tail -2; grep -E "^FAILED|^ERROR|assert .* <=|Error" gpurun_out/t_f8_net.log | head -20
timeout 300 python bench.py --precision f16f8 --steps 10 --warmup 3 --no-cpu-baseline --no-fast > gpurun_out/bench_f8.json 2> gpurun_out/bench_f8.err; echo "bench f16f8 rc $?"; tail -3 gpurun_out/bench_f8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_f8.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "dtype")}, (d.get("e2e") or {}).get("value"), d.get("roofline"), d.get("clocks"))
except Exception as e:
    print("bench unreadable", e)
PY
