# round 2: solver tests after the GMRES / SSOR changes, mitchell_schaefer.ini as written, element-kernel variants on
# the configs[4] workload, SpMV at the headline size
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=short -x -k "gmres or sor or linear_solve or steps or precond or cell" ) > gpurun_out/gpu_tests_h.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/gpu_tests_h.log
timeout 600 python tools/bench_ms.py 7 10 > gpurun_out/bench_ms.log 2>&1; cat gpurun_out/bench_ms.log | tail -4
timeout 600 python tools/bench_ms.py 9 5 >> gpurun_out/bench_ms.log 2>&1; cat gpurun_out/bench_ms.log | tail -3
bash tools/gpu_run_cell.sh "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 vector_gather=true" "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 elem_min_blocks=4" "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 elem_min_blocks=4,vector_gather=true" "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 elem_min_blocks=5,vector_gather=true"
timeout 600 ncu --set full --clock-control none -k regex:k_spmv -s 10 -c 1 -f -o gpurun_out/r02_spmv_256 python bench.py --matrix-free 0 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/ncu_r02_spmv_256.log 2>&1; echo "ncu spmv rc=$?"; tail -2 gpurun_out/ncu_r02_spmv_256.log | cut -c1-600
