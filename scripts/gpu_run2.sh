mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc $?"
cat gpurun_out/bench_ref.json
python scripts/profile_layers.py --precision bf16x3 --out gpurun_out/layers_bf16x3.txt > /dev/null 2> gpurun_out/layers.err
python scripts/profile_layers.py --precision bf16 --out gpurun_out/layers_bf16.txt > /dev/null 2>> gpurun_out/layers.err
python scripts/profile_layers.py --precision fp32 --batch 8 --out gpurun_out/layers_fp32_b8.txt > /dev/null 2>> gpurun_out/layers.err
head -45 gpurun_out/layers_bf16x3.txt; tail -3 gpurun_out/layers.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --no-fast > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 3 -o gpurun_out/prof_conv_tc_r01 \
   python scripts/profile_layers.py --precision bf16x3 --batch 32 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out | head -30
