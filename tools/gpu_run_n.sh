# N ranks on one box: multi-GPU parity tests, then bench.py under torchrun (headline; optional variants)
N=$1; shift
mkdir -p gpurun_out
if [ "$TESTS" != "0" ]; then
( time timeout 1800 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short --maxfail=6 ) > gpurun_out/multi_tests_n$N.log 2>&1; echo "multi tests rc=$?"
tail -12 gpurun_out/multi_tests_n$N.log
fi
run() {
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  tail -1 gpurun_out/bench_n${N}_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('N=$N $name', 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], r.get('host_ms_per_step'))
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_n${N}_$name.err').read()[-2500:])
"
}
for a in "$@"; do
  n=$(echo "$a" | tr ' =,-.' '_____')
  run "$n" $a
done
