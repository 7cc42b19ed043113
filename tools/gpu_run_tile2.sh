# tile kernel: parity tests, launch-shape variants and one ncu --set full capture of the fused apply
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_tile.py -q -m gpu --tb=short -x ) > gpurun_out/tile_tests.log 2>&1; echo "tile tests rc=$?"
tail -15 gpurun_out/tile_tests.log
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
run y4b3
run y6b2 --b200 tile_y=6,tile_min_blocks=2
run y8b1 --b200 tile_y=8,tile_min_blocks=1
run y8b2 --b200 tile_y=8,tile_min_blocks=2
run y4b3lz64 --b200 tile_lz=64
run old --b200 tile=false
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_tile_apply -s 20 -c 1 -o gpurun_out/tile_apply_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 > gpurun_out/ncu_tile.log 2>&1; echo "ncu rc=$?"
