#!/bin/bash
# One-shot validation on a B200 box (what the round-end driver runs, plus the profiling passes):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -3 gpurun_out/t_all_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --mode full --steps 5 --warmup 3 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err; echo "full rc $?"; tail -2 gpurun_out/bench_full_n1.err
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; echo "train rc $?"
echo "t=$(( $(date +%s) - T0 ))s"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --metrics $M --clock-control none -k regex:conv_tc_kernel -s 179 -c 179 --csv --log-file gpurun_out/conv_traffic.csv \
    python scripts/ncu_conv_step.py --precision f16f8 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc $?"; tail -2 gpurun_out/ncu_traffic.log
gzip -f gpurun_out/conv_traffic.csv
echo "t=$(( $(date +%s) - T0 ))s"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc $?"
gzip -f gpurun_out/launches_bench.csv
echo "t=$(( $(date +%s) - T0 ))s"
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref", "bench_full_n1", "bench_train_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, (d.get("e2e") or {}).get("value"), (d.get("e2e_u8_input") or {}).get("value"),
              (d.get("roofline") or {}).get("frac"), (d.get("alt_parity_mode") or {}).get("value"), (d.get("fast_mode") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
