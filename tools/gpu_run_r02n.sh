# same-box A/B of the structured apply: commit f6769d9 against the current tree.  Needs the worktree first:
#   git worktree add ab_old f6769d9 && (cd ab_old && python -c "import sys; sys.path.insert(0, \".\"); from dune_copasi_b200 import build as B; B.build(); import bench; bench.precompile()")
mkdir -p gpurun_out
run() {
  name=$1; dir=$2; shift 2
  ( cd $dir && timeout 600 python bench.py --no-cpu-baseline --no-assembled --no-e2e --steps 5 --warmup 3 "$@" ) > gpurun_out/bench_ab2_$name.json 2> gpurun_out/bench_ab2_$name.err
  tail -1 gpurun_out/bench_ab2_$name.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']; q=d.get('q1_variant')
print('$name: ms/step %.2f' % d['ms_per_step'], r['kernel'], round(r['avg_launch_ms'],4), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['clocks']['power_w'], 'Q1', q and round(q['ms_per_step'],2), q and round(q['roofline']['avg_launch_ms'],4))" || tail -5 gpurun_out/bench_ab2_$name.err
}
Y=model.time_step_operator.linear_solver.b200.yfree
run old ab_old
run new_default .
run new_noyfree . --set $Y=false
run new_noyfree_nomask . --set $Y=false --b200 struct_nomask=true
run new_hostvol . --b200 host_vol=true
run new_noyfree_nomask_hostvol . --set $Y=false --b200 struct_nomask=true,host_vol=true
