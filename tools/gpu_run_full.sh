# the round-end check: GPU suite, smoke, optionally the contract bench and the reference arm
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu --tb=short "${PYTEST_ARGS:--x}" ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
if [ "$1" = "bench" ]; then
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
fi
