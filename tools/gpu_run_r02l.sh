# round 2: packed connectivity of the element kernels + fix of the scaled apply: suite, cell workloads
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/gpu_tests.log
bash tools/gpu_run_cell.sh "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05" "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 packed_conn=false" "--workload cell --cells 96 --steps 3 --warmup 2 --dt 0.05"
