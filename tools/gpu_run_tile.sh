# first GPU run of the tile-marching kernels: parity tests, then the headline bench with the tile
# drivers / fused BiCGSTAB on and off (one summary line each)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_tile.py -q -m gpu --tb=short -x "$@" ) > gpurun_out/tile_tests.log 2>&1; echo "tile tests rc=$?"
tail -40 gpurun_out/tile_tests.log
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-q1 "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -1 gpurun_out/bench_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$name', 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], 'launches', d['gpu_launches'], 'top', r['kernel'], round(r['avg_launch_ms'],4))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_$name.err').read()[-1500:])
"
}
run tile
run tile_unfused --set model.time_step_operator.linear_solver.b200.fused=false
run old --b200 tile=false
