mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
if [ "$1" = "bench" ]; then
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
fi
for wl in cell cell10; do
timeout 300 python bench.py --workload $wl --cells 64 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl', 'dofs', d['dofs'], 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, d['solver_stats']['linear_half_iterations'])"
done
