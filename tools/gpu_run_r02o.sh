# round 2: mask-free element apply on the configs[4] workload (same box), ncu of the final structured apply
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu --tb=short -x -k "cell or two_disks or advection or numerical" ) 2>&1 | tail -2
bash tools/gpu_run_cell.sh "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05" "--workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05 --b200 struct_nomask=false" "--workload cell --cells 96 --steps 3 --warmup 2 --dt 0.05"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_struct_apply -s 20 -c 1 -f -o gpurun_out/r02_struct_apply_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/ncu_r02_struct_apply_final.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/launches_r02_final.log 2>&1; echo "launch list rc=$?"
