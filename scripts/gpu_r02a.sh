#!/bin/bash
# Round-2 first GPU pass: full GPU suite (incl. the headline-config parity tests), smoke, the default bench line (now with
# parity / cpu_cfg1 / roofline_aux / train_step / full_pipeline), the TMA stride probe, NMS v1-vs-v2 timing, compute-sanitizer.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -5 gpurun_out/t_all_gpu.log | cut -c1-400
grep -E "^(FAILED|ERROR)" gpurun_out/t_all_gpu.log | cut -c1-250 | head -40
grep -E "R101 .* 480x640|free-running|PRN (small|prod)" gpurun_out/t_all_gpu.log | head -20
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"; tail -3 gpurun_out/bench_n1.err | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_stride_probe scripts/exp/tma_stride_probe.cu 2> /dev/null
timeout 120 gpurun_out/tma_stride_probe > gpurun_out/tma_stride_probe.txt 2>&1; echo "probe rc $?"; rm -f gpurun_out/tma_stride_probe
MPN_NMS_V2=0 timeout 300 python - > gpurun_out/nms_v1_stages.txt 2>&1 <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from multiposenet.pytorch_b200 import ops, synthetic
c3, b3 = synthetic.cfg3_detections(32, seed=3)
c3, b3 = torch.from_numpy(c3).cuda(), torch.from_numpy(b3).cuda()
st = []
for _ in range(4):
    ops.filter_sort_nms(c3, b3, 0.05, 0.5, max_cand=4224, stage_ms=st)
print("v1 kernels (MPN_NMS_V2=0), cfg3 feed, batch 32: filter, sort, gather, mask, reduce ms =", np.median(np.array(st), axis=0).round(4).tolist())
PY
cat gpurun_out/nms_v1_stages.txt | tail -1
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_target.py --train > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/sanitizer_memcheck.log | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/sanitizer_racecheck.log | cut -c1-300
echo "t=$(( $(date +%s) - T0 ))s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print("bench", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"),
          "parity", d.get("parity"), "clocks", d.get("clocks"))
    for k in ("train_step", "full_pipeline", "cpu_cfg1"):
        v = d.get(k) or {}
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error", "cpu_ms", "gpu_ms_e2e", "speedup_e2e")})
    for r in (d.get("roofline_aux") or {}).get("kernels", []):
        print("  aux %-34s %-60s %.4f ms %.0f GB/s" % (r["regime"][:34], r["kernel"][:60], r["ms"], r["achieved"] or 0))
    print((d.get("roofline_aux") or {}).get("error"))
except Exception as e:
    print("bench unreadable", e)
PY
