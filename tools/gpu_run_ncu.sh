# one ncu --set full capture: gpu_run_ncu.sh <name> <kernel regex> <skip> <bench args...>
mkdir -p gpurun_out
name=$1; regex=$2; skip=$3; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/$name python bench.py --no-cpu-baseline --no-e2e --no-q1 --no-assembled "$@" > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
