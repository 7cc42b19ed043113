# round 2, 8 ranks: a few many-rank parity cases, then the headline and the configs[4] workload with device timelines
N=${1:-8}; shift
mkdir -p gpurun_out
if [ "$TESTS" != "0" ]; then ( time timeout 900 python -m pytest "tests/test_gpu_multi.py::test_many_rank_time_stepping_matches_serial_oracle[grayscott3d-1--8]" "tests/test_gpu_multi.py::test_many_rank_time_stepping_matches_serial_oracle[cell10_nested-1--8]" -q -m gpu --tb=short ) > gpurun_out/multi_tests_n$N.log 2>&1; echo "multi tests rc=$?"; fi
tail -6 gpurun_out/multi_tests_n$N.log
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline --timeline gpurun_out/timeline_n${N}_$name.json "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  tail -1 gpurun_out/bench_n${N}_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('N=$N $name', 'dofs', d['dofs'], 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], d.get('timeline'), 'launches', d['gpu_launches'], 'setup_s', round(d['setup_s'],1))
    t = json.load(open('gpurun_out/timeline_n${N}_$name.json'))
    for k, v in list(t['kernels'].items())[:12]: print('     ', k, v)
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_n${N}_$name.err').read()[-2500:])
"
}
for a in "$@"; do
  n=$(echo "$a" | tr ' =,-.' '_____')
  run "$n" $a
done
