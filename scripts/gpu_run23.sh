mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "exit full gpu suite: $?"
tail -5 gpurun_out/t_all_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit smoke: $?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
cat gpurun_out/bench_n1.json | cut -c1-2000
