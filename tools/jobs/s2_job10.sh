set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_g.log 2>&1; tail -5 gpurun_out/s2_pytest_g.log
