#!/bin/bash
# Round-end validation + evidence on one B200:  gpurun --timeout 2700 -- 'bash scripts/gpu_final.sh'
mkdir -p gpurun_out; rm -f gpurun_out/headline_parity.jsonl
T0=$(date +%s)
timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"; tail -3 gpurun_out/t_all_gpu.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/t_all_gpu.log | cut -c1-250 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
echo "t=$(( $(date +%s) - T0 ))s"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/bench_n1.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python scripts/profile_layers.py --precision f16f8 --out gpurun_out/layers_f16f8.txt > /dev/null 2> gpurun_out/lay.err; echo "layers rc $?"
NL=$(python scripts/ncu_conv_step.py --precision f16f8 --steps 1 --names-out gpurun_out/conv_launch_names.txt 2>/dev/null | grep -o "[0-9]* conv launches" | head -1 | grep -o "^[0-9]*")
echo "conv launches per step: $NL"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 500 ncu --metrics $M --clock-control none -k regex:conv_tc_kernel -s $NL -c $NL --csv --log-file gpurun_out/conv_traffic.csv \
    python scripts/ncu_conv_step.py --precision f16f8 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc $?"; tail -2 gpurun_out/ncu_traffic.log
gzip -f gpurun_out/conv_traffic.csv
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast --no-extras > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list rc $?"
gzip -f gpurun_out/launches_bench.csv
echo "t=$(( $(date +%s) - T0 ))s"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_target.py --train > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -2 gpurun_out/sanitizer_memcheck.log | cut -c1-200
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; tail -2 gpurun_out/sanitizer_racecheck.log | cut -c1-200
grep -o "in mpn_[a-z_]*\.cuh\?:[0-9]*" gpurun_out/sanitizer_racecheck.log | sort | uniq -c
echo "t=$(( $(date +%s) - T0 ))s"
cat gpurun_out/headline_parity.jsonl
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", (d.get("e2e") or {}).get("value"), "u8", (d.get("e2e_u8_input") or {}).get("value"),
              "frac", r.get("frac"), "pipe", r.get("tensor_pipe_frac"), "passes", r.get("mma_passes"), "alt", (d.get("alt_parity_mode") or {}).get("value"),
              "fast", (d.get("fast_mode") or {}).get("value"), d.get("clocks"))
        print("   parity", d.get("parity"))
        for k in ("train_step", "full_pipeline", "cpu_cfg1"):
            v = d.get(k) or {}
            print("  ", k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error", "cpu_ms", "gpu_ms_e2e", "speedup_e2e") if v.get(kk) is not None})
    except Exception as e:
        print(f, "unreadable", e)
PY
