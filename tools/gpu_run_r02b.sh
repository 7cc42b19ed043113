# round 2: x-pair pre-reduced scatter of the per-cell driver -- parity first, then the headline and Q1 numbers, the RED micro-benchmark
mkdir -p gpurun_out
./tools/micro/red_sectors > gpurun_out/red_sectors.log 2>&1; cat gpurun_out/red_sectors.log
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_q1.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -x -k "residual or apply or diag or q1 or full" ) > gpurun_out/gpu_tests_b.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/gpu_tests_b.log
( time timeout 900 python bench.py --no-cpu-baseline --no-assembled --no-e2e ) > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02b.json').read().strip().splitlines()[-1]); r = d['roofline']
print('P1 ms/step %.2f profiled %.2f' % (d['ms_per_step'], r['profiled_ms_per_step']), r['kernel'], r['avg_launch_ms'], r['breakdown_ms_per_step'])
q = d['q1_variant']; print('Q1 ms/step %.2f' % q['ms_per_step'], q['roofline']['kernel'], q['roofline']['avg_launch_ms'], q['roofline']['breakdown_ms_per_step'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_struct_apply -s 20 -c 1 -f -o gpurun_out/r02_struct_apply_xpair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/ncu_r02_struct_apply_xpair.log 2>&1; echo "ncu rc=$?"
