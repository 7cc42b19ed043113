# round 2: group-transposed scatter in the CSR fill / block diagonal -- parity, then the matrix-based bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu --tb=short -x -k "csr or block_diag or linear_solve or steps or numerical" ) 2>&1 | tail -3
for c in 128 160; do
timeout 600 python bench.py --cells $c --matrix-free 0 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-q1 --no-assembled > gpurun_out/bench_asm_$c.json 2> gpurun_out/bench_asm_$c.err
tail -1 gpurun_out/bench_asm_$c.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('assembled $c^3: ms/step %.2f value %.3e' % (d['ms_per_step'], d['value']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()}, 'linearizations', d['solver_stats']['linearizations'], 'steps', d['steps'])"
done
