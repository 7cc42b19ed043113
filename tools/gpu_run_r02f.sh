# round 2, 2 ranks: new single-GPU cases (tables, pinned Poisson, Q1 reduce), general peer halo on the RCB partition
N=2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_q1.py -q -m gpu --tb=short --maxfail=5 -k "tables or poisson_pinned or reduce" ) > gpurun_out/gpu_tests_f.log 2>&1; echo "single tests rc=$?"; tail -5 gpurun_out/gpu_tests_f.log
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short --maxfail=4 -k "two_rank_time and (cell10_nested or tables or cell3d or two_disks or advection)" ) > gpurun_out/multi_tests_f.log 2>&1; echo "multi tests rc=$?"; tail -6 gpurun_out/multi_tests_f.log
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline --timeline gpurun_out/timeline_n${N}_$name.json "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  tail -1 gpurun_out/bench_n${N}_$name.json | python -c "
import sys, json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('N=$N $name', 'dofs', d['dofs'], 'ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'halfits', d['solver_stats']['linear_half_iterations'], d.get('timeline'), 'launches', d['gpu_launches'], 'setup_s', round(d['setup_s'],1))
    t = json.load(open('gpurun_out/timeline_n${N}_$name.json'))
    for k, v in list(t['kernels'].items())[:10]: print('     ', k, v)
except Exception as e:
    print('$name failed', e); print(open('gpurun_out/bench_n${N}_$name.err').read()[-2500:])
"
}
run cell10 --workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05
DCB_PEER_GENERAL_HALO=0 run cell10_nccl_halo --workload cell10 --mesh nested --cells 96 --steps 3 --warmup 2 --dt 0.05
