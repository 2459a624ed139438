#!/bin/bash
# Quick GPU pass: the full GPU suite and the default bench line.   gpurun --timeout 1500 -- 'bash scripts/gpu_quick.sh [pytest -k expr]'
mkdir -p gpurun_out
T0=$(date +%s)
if [ -n "$1" ]; then
  timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider -k "$1" > gpurun_out/t_gpu.log 2>&1; echo "exit gpu tests (-k $1): $?"
else
  timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "exit gpu suite: $?"
fi
tail -3 gpurun_out/t_gpu.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | cut -c1-250 | head -40
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"; tail -3 gpurun_out/bench_n1.err | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
# same-box A/B: AB_ENV="MPN_NO_H8=0;MPN_UP_TMA=0" runs one short bench per setting (device-resident value only)
if [ -n "$AB_ENV" ]; then
  IFS=';' read -ra SETS <<< "$AB_ENV"
  for S in "base" "${SETS[@]}"; do
    if [ "$S" = "base" ]; then E=""; else E="$S"; fi
    env $E timeout 300 python bench.py --no-extras --no-cpu-baseline --no-fast --steps 20 > gpurun_out/ab.json 2> gpurun_out/ab.err
    python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print('AB %-28s value %.1f img/s  ms %.3f  conv_ms %.3f  frac %.4f  pipe %.3f  sm %s' % ('$S', d['value'], d['ms_per_step'], r.get('kernel_ms_per_step',0), r.get('frac',0), r.get('tensor_pipe_frac',0), (d.get('clocks') or {}).get('sm_mhz')))
except Exception as e:
    print('AB $S failed', e); print(open('gpurun_out/ab.err').read()[-600:])
"
  done
  echo "t=$(( $(date +%s) - T0 ))s"
fi
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("bench", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", (d.get("e2e") or {}).get("value"), "u8", (d.get("e2e_u8_input") or {}).get("value"),
          "frac", r.get("frac"), "conv_ms", r.get("kernel_ms_per_step"), "alt", (d.get("alt_parity_mode") or {}).get("value"), "fast", (d.get("fast_mode") or {}).get("value"))
    print("parity", d.get("parity"))
    print("clocks", d.get("clocks"))
    for k in ("train_step", "full_pipeline", "cpu_cfg1"):
        v = d.get(k) or {}
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error", "cpu_ms", "gpu_ms_e2e", "speedup_e2e", "allreduce_alone_ms") if v.get(kk) is not None})
    for r in (d.get("roofline_aux") or {}).get("kernels", []):
        print("  aux %-28s %-44s %.4f ms %.0f GB/s" % (r["regime"][:28], r["kernel"][:44], r["ms"], r["achieved"] or 0))
    if (d.get("roofline_aux") or {}).get("error"):
        print(d["roofline_aux"])
except Exception as e:
    print("bench unreadable", e)
PY
