mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -p no:cacheprovider -k "uint8 or train or golden" > gpurun_out/t_net.log 2>&1; echo "exit: $?"; tail -4 gpurun_out/t_net.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:(d[k] if not isinstance(d[k],dict) else {kk:d[k][kk] for kk in list(d[k])[:4]}) for k in ("value","e2e","e2e_u8_input","ms_per_step")})
PY
