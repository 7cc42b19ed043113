#!/bin/bash
# usage: tools/sweep.sh CELLS "variant1" "variant2" ...   (variant = --b200 string)
cells=$1; shift
for v in "$@"; do
  echo "== $v"
  timeout 300 python bench.py --cells $cells --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-q1 --b200 "$v" 2>&1 | tail -1 | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); r=d['roofline']
    print('ms/step %.2f  value %.3e  top %s avg %.3f ms frac %.3f' % (d['ms_per_step'], d['value'], r['kernel'], r['avg_launch_ms'], r['frac']), {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()})
except Exception as e:
    print('ERR', l[-400:])
"
done
