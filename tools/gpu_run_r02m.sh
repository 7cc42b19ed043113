# round 2: A/B on ONE box (boxes differ by ~8 % under the power cap): prefetch, D^-1 in the apply, launch shape
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --no-cpu-baseline --no-assembled --no-q1 --no-e2e --steps 5 --warmup 3 "$@" > gpurun_out/bench_ab_$name.json 2> gpurun_out/bench_ab_$name.err
  tail -1 gpurun_out/bench_ab_$name.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name: ms/step %.2f' % d['ms_per_step'], r['kernel'], round(r['avg_launch_ms'],4), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['clocks']['power_w'])"
}
Y=model.time_step_operator.linear_solver.b200.yfree
run default
run noprefetch --b200 struct_prefetch=false
run noyfree --set $Y=false
run neither --b200 struct_prefetch=false --set $Y=false
run default_again
run minb13 --b200 struct_min_blocks=13
run minb11 --b200 struct_min_blocks=11
