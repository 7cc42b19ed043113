mkdir -p gpurun_out
for lv in 5 7; do timeout 300 python tools/bench_ms.py $lv 5 2>&1 | tail -4; done > gpurun_out/bench_ms.log 2>&1; cat gpurun_out/bench_ms.log
MS_DT=1.0 timeout 300 python tools/bench_ms.py 7 5 2>&1 | tail -4 | tee -a gpurun_out/bench_ms.log
