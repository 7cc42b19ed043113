mkdir -p gpurun_out
for el in q1 p1; do
for v in "struct_march_apply=0" "struct_march_apply=8" "struct_march_apply=16" "struct_march_apply=8,struct_min_blocks=10"; do
  echo "== $el $v"
  timeout 300 python bench.py --element $el --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-q1 --b200 "$v" 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); r=d['roofline']
    print('ms/step %.2f  value %.3e  top %s avg %.3f ms frac %.3f' % (d['ms_per_step'], d['value'], r['kernel'], r['avg_launch_ms'], r['frac']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()})
except Exception as e:
    print('ERR', l[-400:])
"
done; done
