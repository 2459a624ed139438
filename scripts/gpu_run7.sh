mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_network.log 2>&1
echo "exit network: $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/t_network.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
cat gpurun_out/bench_ref.json
