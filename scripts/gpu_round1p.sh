#!/bin/bash
# Validation + measurement call for the r01p build (new: PRN assignment, evaluate pipeline, full-pipeline bench, DRAM traffic pass).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round1p.sh'
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -5 gpurun_out/t_all_gpu.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
timeout 420 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --mode full --steps 5 --warmup 3 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err; echo "full rc $?"; tail -2 gpurun_out/bench_full_n1.err
echo "t=$(( $(date +%s) - T0 ))s"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --metrics $M --clock-control none -k regex:conv_tc_kernel -s 179 -c 179 --csv --log-file gpurun_out/conv_traffic.csv \
    python scripts/ncu_conv_step.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc $?"; tail -2 gpurun_out/ncu_traffic.log
echo "t=$(( $(date +%s) - T0 ))s"
gzip -f gpurun_out/conv_traffic.csv
python - <<'PY'
import json
for f in ("bench_n1", "bench_full_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("config", {}).get("persons_per_image"))
    except Exception as e:
        print(f, "unreadable", e)
PY
