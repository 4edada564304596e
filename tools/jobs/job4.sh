set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2_pytest_sharded.log 2>&1; tail -30 gpurun_out/r2_pytest_sharded.log
