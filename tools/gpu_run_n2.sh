mkdir -p gpurun_out
N=${1:-2}
# smoke first, bounded: a hang in the peer-memory kernels must not eat the budget
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 tools/mgpu_check.py grayscott3d 2 1 > gpurun_out/peer_smoke.log 2>&1; rc=$?
echo "peer smoke rc=$rc"; grep -E "mgpu_check|Error|error" gpurun_out/peer_smoke.log | head -5
if [ $rc -ne 0 ]; then tail -30 gpurun_out/peer_smoke.log; exit 1; fi
if [ "$N" = "2" ]; then
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -5 gpurun_out/multi_tests.log
fi
for el in ${2:-p1 q1}; do
for peer in 1 0; do
DCB_PEER_COLLECTIVES=$peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --element $el --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-q1 > gpurun_out/bench_${el}_n${N}_peer${peer}.json 2> gpurun_out/bench_${el}_n${N}_peer${peer}.err; echo "bench $el n$N peer=$peer rc=$?"
tail -1 gpurun_out/bench_${el}_n${N}_peer${peer}.json | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.2f  value %.3e' % (d['ms_per_step'], d['value']), d['config']['collectives'][:20], {k: round(v,2) for k,v in r['breakdown_ms_per_step'].items()})"
done; done
