mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for g in detect conv network; do
  timeout 900 python -m pytest tests/test_gpu_$g.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_$g.log 2>&1
  echo "exit $g: $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -15 gpurun_out/t_conv.log; tail -15 gpurun_out/t_network.log
python scripts/profile_layers.py --precision bf16x3 --out gpurun_out/layers_bf16x3.txt > /dev/null 2> gpurun_out/layers.err
python scripts/profile_layers.py --precision bf16 --out gpurun_out/layers_bf16.txt > /dev/null 2>> gpurun_out/layers.err
head -32 gpurun_out/layers_bf16x3.txt; head -14 gpurun_out/layers_bf16.txt; tail -3 gpurun_out/layers.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --steps 10 --warmup 3 --streams 0 --no-cpu-baseline > gpurun_out/bench_n1_nostreams.json 2> gpurun_out/bench_n1_nostreams.err; echo "bench nostreams rc $?"
cat gpurun_out/bench_n1_nostreams.json
