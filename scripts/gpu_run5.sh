mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for g in conv network; do
  timeout 900 python -m pytest tests/test_gpu_$g.py -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/t_$g.log 2>&1
  echo "exit $g: $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -8 gpurun_out/t_conv.log; tail -8 gpurun_out/t_network.log
python scripts/profile_layers.py --precision bf16x3 --out gpurun_out/layers_bf16x3.txt > /dev/null 2> gpurun_out/layers.err
MPN_SPLIT_BN256=1 python scripts/profile_layers.py --precision bf16x3 --out gpurun_out/layers_bf16x3_bn256.txt > /dev/null 2>> gpurun_out/layers.err
python scripts/profile_layers.py --precision bf16 --out gpurun_out/layers_bf16.txt > /dev/null 2>> gpurun_out/layers.err
head -24 gpurun_out/layers_bf16x3.txt; head -24 gpurun_out/layers_bf16x3_bn256.txt; tail -3 gpurun_out/layers.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv \
   python scripts/profile_layers.py --precision bf16x3 --batch 32 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 44 -c 6 -o gpurun_out/prof_conv_tc_r01c \
   python scripts/profile_layers.py --precision bf16x3 --batch 32 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
