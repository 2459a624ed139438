#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -p no:cacheprovider > gpurun_out/t_conv.log 2>&1; echo "conv rc $?"; tail -1 gpurun_out/t_conv.log
timeout 200 python scripts/profile_layers.py --precision f16f8 > gpurun_out/layers_f16f8_v.txt 2> gpurun_out/lay.err; head -12 gpurun_out/layers_f16f8_v.txt
timeout 200 python scripts/profile_layers.py --precision bf16x3 > gpurun_out/layers_bf16x3_v.txt 2>> gpurun_out/lay.err; head -8 gpurun_out/layers_bf16x3_v.txt
timeout 300 python bench.py --precision f16f8 --no-cpu-baseline --no-fast --steps 10 > gpurun_out/bench_f8_v.json 2> gpurun_out/bench_f8.err; echo "bench f8 rc $?"
python - <<'PY'
import json
for f in ("bench_f8_v",):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
PY
