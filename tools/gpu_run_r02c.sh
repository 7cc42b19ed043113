# round 2: whole GPU suite + smoke + the headline line after the fused-link sweeps
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -12 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py --no-cpu-baseline --no-q1 --no-assembled --timeline gpurun_out/timeline_n1_links.json ) > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02c.json').read().strip().splitlines()[-1]); r = d['roofline']
print('ms/step %.2f e2e %.2f profiled %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['profiled_ms_per_step']), r['breakdown_ms_per_step'], d.get('timeline'), d['gpu_launches'])
PY
