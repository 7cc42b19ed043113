mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "jacobian_csr or time_steps or linear_solve or gmres" > gpurun_out/csr_tests.log 2>&1; echo "csr tests rc=$?"
tail -5 gpurun_out/csr_tests.log
for v in gather scatter; do
for wl in "--cells 128" "--workload cell --cells 64"; do
echo "== $v $wl"
timeout 300 python bench.py $wl --matrix-free 0 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-q1 --b200 "csr_fill=$v" 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, d['solver_stats']['linearizations'])"
done; done
