# configs[4]: the 10-species cell model on the nested-compartment tetrahedral mesh (one summary line per run)
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 900 python bench.py --no-cpu-baseline --no-q1 --no-assembled "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -1 gpurun_out/bench_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$name', 'dofs', d['dofs'], 'elements', d.get('elements'), 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], 'top', r['kernel'], round(r['avg_launch_ms'],4), 'frac', round(r['frac'],3), 'setup_s', round(d['setup_s'],1))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_$name.err').read()[-1500:])
"
}
for a in "$@"; do
  n=$(echo "$a" | tr ' =,-' '____')
  run "$n" $a
done
