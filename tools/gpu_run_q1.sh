mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_q1.py -x -q > gpurun_out/q1_tests.log 2>&1; echo "q1 tests rc=$?" 
tail -5 gpurun_out/q1_tests.log
timeout 300 python bench.py --element q1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench q1 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p1.json 2> gpurun_out/bench_p1.err; echo "bench p1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_q1_apply -s 20 -c 1 -o gpurun_out/q1_apply python bench.py --element q1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_q1.log 2>&1; echo "ncu rc=$?"
cut -c1-1500 gpurun_out/bench_q1.json
