# round 2: transposed group scatter of the element kernels -- whole GPU suite, cell workloads, ncu of the apply kernel
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/gpu_tests.log
bash tools/gpu_run_cell.sh "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05" "--workload cell --cells 96 --steps 3 --warmup 2 --dt 0.05"
bash tools/gpu_run_ncu.sh r02_elem_apply_nested96 dc_k_jacobian_apply_volume_1 10 --workload cell10 --mesh nested --cells 96 --steps 1 --warmup 1 --dt 0.05
