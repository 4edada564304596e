set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_j.log 2>&1; tail -5 gpurun_out/s2_pytest_j.log
timeout 600 python tools/jit_bench.py 16 > gpurun_out/s2_jit_bench2.jsonl 2>&1; cat gpurun_out/s2_jit_bench2.jsonl
