mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-q1 "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -1 gpurun_out/bench_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$name', 'ms/step %.2f' % d['ms_per_step'], {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], 'top', r['kernel'], round(r['avg_launch_ms'],4))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_$name.err').read()[-1500:])
"
}
for v in "$@"; do run "$(echo $v | tr ',=' '__')" --b200 $v; done
