mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for g in detect conv network; do
  timeout 900 python -m pytest tests/test_gpu_$g.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_$g.log 2>&1
  echo "exit $g: $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -6 gpurun_out/t_detect.log; tail -6 gpurun_out/t_conv.log; tail -6 gpurun_out/t_network.log
python scripts/profile_layers.py --precision bf16x3 --out gpurun_out/layers_bf16x3.txt > /dev/null 2> gpurun_out/layers.err
head -16 gpurun_out/layers_bf16x3.txt; tail -3 gpurun_out/layers.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc $?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc $?"
cat gpurun_out/bench_ref_n2.json
