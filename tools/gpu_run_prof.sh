mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_p1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 > gpurun_out/launches_p1.log 2>&1; echo "launch list p1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_q1.csv python bench.py --element q1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_q1.log 2>&1; echo "launch list q1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_struct_apply -s 20 -c 1 -o gpurun_out/p1_apply python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 > gpurun_out/ncu_p1.log 2>&1; echo "ncu p1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_q1_apply -s 20 -c 1 -o gpurun_out/q1_apply2 python bench.py --element q1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_q1b.log 2>&1; echo "ncu q1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_q1_march_residual -s 2 -c 1 -o gpurun_out/q1_march_res python bench.py --element q1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_q1c.log 2>&1; echo "ncu q1 march rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:k_halo\|k_allreduce -c 2 -o gpurun_out/none python -c "print('skip')" > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
