#!/bin/bash
# Measurement variants next to the headline: python bench.py with other schemes / solvers /
# workloads, one summary line each.  Usage: bash tools/bench_variants.sh [name ...]
run() { name=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-q1 "$@" > gpurun_out/var_$name.log 2>&1; tail -1 gpurun_out/var_$name.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
    print('$name', 'ms/step=%.2f'%d['ms_per_step'], 'DOFupd/s=%.3e'%d['value'], 'e2e=%.3e'%d['e2e']['value'], 'dofs=%d'%d['dofs'], 'top=%s frac=%.3f'%(r.get('kernel'), r.get('frac',0)), {k: round(v,2) for k,v in (r.get('breakdown_ms_per_step') or {}).items()}, 'halfits=%d'%d['solver_stats']['linear_half_iterations'], 'launches=%d'%d['gpu_launches'])
except Exception as e:
    print('$name FAILED', e)
"; }
want() { [ $# -eq 0 ] && return 0; for w in "$@"; do [ "$w" == "$NAME" ] && return 0; done; return 1; }
mkdir -p gpurun_out
NAME=headline;         want "$@" && run $NAME
NAME=implicit_euler;   want "$@" && run $NAME --rk ImplicitEuler
NAME=blockjacobi;      want "$@" && run $NAME --prec BlockJacobi
NAME=s2d_4096;         want "$@" && run $NAME --dim 2 --cells 4096
NAME=s2d_1024;         want "$@" && run $NAME --dim 2 --cells 1024
NAME=matrix_based_128; want "$@" && run $NAME --cells 128 --matrix-free 0
NAME=atomic_128;       want "$@" && run $NAME --cells 128 --scheme atomic
NAME=cell_96;          want "$@" && run $NAME --workload cell --cells 96
true
