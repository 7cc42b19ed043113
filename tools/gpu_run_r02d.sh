# round 2, N ranks: bounded smoke of the fused-link collectives, multi-GPU parity tests, bench with the device timeline
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 tools/mgpu_check.py grayscott3d 2 1 > gpurun_out/peer_smoke.log 2>&1; rc=$?
echo "peer smoke rc=$rc"; grep -E "mgpu_check|Error|error" gpurun_out/peer_smoke.log | head -5
if [ $rc -ne 0 ]; then tail -30 gpurun_out/peer_smoke.log; exit 1; fi
if [ "$TESTS" != "0" ]; then
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short --maxfail=4 ) > gpurun_out/multi_tests_n$N.log 2>&1; echo "multi tests rc=$?"
tail -6 gpurun_out/multi_tests_n$N.log
fi
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline --timeline gpurun_out/timeline_n${N}_$name.json "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  tail -1 gpurun_out/bench_n${N}_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('N=$N $name', 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], d.get('timeline'), 'launches', d['gpu_launches'])
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
