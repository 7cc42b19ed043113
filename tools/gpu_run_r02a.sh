# round 2: headline bench with the device timeline, the launch list and one --set full capture of the apply kernel
mkdir -p gpurun_out
( time timeout 900 python bench.py --no-cpu-baseline --no-q1 --no-assembled --timeline gpurun_out/timeline_n1.json ) > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02a.json').read().strip().splitlines()[-1]); r = d['roofline']
print('ms/step %.2f e2e %.2f profiled %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['profiled_ms_per_step']), r['breakdown_ms_per_step'], d.get('timeline'))
t = json.load(open('gpurun_out/timeline_n1.json'))
for k, v in list(t['kernels'].items())[:14]: print('  ', k, v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/r02_launches_bench256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/launches_r02.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dc_k_struct_apply -s 20 -c 1 -f -o gpurun_out/r02_struct_apply python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/ncu_r02_struct_apply.log 2>&1; echo "ncu rc=$?"
