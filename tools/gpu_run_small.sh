mkdir -p gpurun_out
for v in "" "model.time_step_operator.linear_solver.b200.speculation=false"; do
for c in 128 256; do
echo "== cells $c set=$v"
timeout 300 python bench.py --cells $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-q1 --set "$v" 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'sum %.2f' % sum(r['breakdown_ms_per_step'].values()), {k: round(v,2) for k,v in r['host_ms_per_step'].items()}, d['solver_stats']['linear_half_iterations'])"
done; done
