# round 2: CSR fill with transposed runs, BiCGSTAB without the stored preconditioned vectors: whole suite, then benches
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/gpu_tests.log
for c in 128; do
timeout 600 python bench.py --cells $c --matrix-free 0 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/bench_asm_$c.json 2> gpurun_out/bench_asm_$c.err
tail -1 gpurun_out/bench_asm_$c.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('assembled $c^3: ms/step %.2f value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'linearizations', d['solver_stats']['linearizations'], 'steps', d['steps'])"
done
for y in true false; do
timeout 600 python bench.py --no-cpu-baseline --no-assembled --set model.time_step_operator.linear_solver.b200.yfree=$y > gpurun_out/bench_yfree_$y.json 2> gpurun_out/bench_yfree_$y.err
tail -1 gpurun_out/bench_yfree_$y.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']; q=d.get('q1_variant')
print('yfree=$y: ms/step %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']), r['kernel'], round(r['avg_launch_ms'],4), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], 'Q1 ms/step', q and round(q['ms_per_step'],2), q and q['roofline']['avg_launch_ms'])"
done
