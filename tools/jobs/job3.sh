set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_b.log 2>&1; tail -5 gpurun_out/r2_pytest_b.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 300 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
