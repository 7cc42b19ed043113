mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_q1.py -x -q -k marching > gpurun_out/march_tests.log 2>&1; echo "march tests rc=$?"
tail -5 gpurun_out/march_tests.log
for el in q1 p1; do
for v in "struct_min_blocks=7" "struct_min_blocks=8" "struct_min_blocks=10" "struct_threads=32,struct_min_blocks=12" "struct_threads=32,struct_min_blocks=16" "struct_threads=96,struct_min_blocks=4" "struct_threads=128,struct_min_blocks=3"; do
  echo "== $el $v"
  timeout 300 python bench.py --element $el --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-q1 --b200 "struct_march=0,$v" 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); r=d['roofline']
    print('ms/step %.2f  value %.3e  top %s avg %.3f ms frac %.3f' % (d['ms_per_step'], d['value'], r['kernel'], r['avg_launch_ms'], r['frac']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()})
except Exception as e:
    print('ERR', l[-400:])
"
done; done
