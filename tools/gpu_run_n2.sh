mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -5 gpurun_out/multi_tests.log
for el in p1 q1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --element $el --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${el}_n2.json 2> gpurun_out/bench_${el}_n2.err; echo "bench $el n2 rc=$?"
tail -1 gpurun_out/bench_${el}_n2.json | cut -c1-400
done
