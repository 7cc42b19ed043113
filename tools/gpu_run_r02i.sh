mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu --tb=short -x -k "sor or gmres or precond" ) 2>&1 | tail -3
for g in true false; do
echo "sor_graph=$g"
MS_DT=1.0 MS_SET="model.time_step_operator.linear_solver.b200.sor_graph=$g" timeout 300 python tools/bench_ms.py 7 5 2>&1 | grep "sor_sweep=False" | tee -a gpurun_out/bench_ms.log
done
