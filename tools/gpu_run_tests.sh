mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x "$@" ) > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/gpu_tests.log
