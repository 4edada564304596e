set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_a.log 2>&1; tail -5 gpurun_out/r2_pytest_a.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 1500 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ntt2_pass" -c 6 -o gpurun_out/r2_ntt2_a python tools/prof_prove.py 20 4 1 dev 1 > gpurun_out/r2_ntt2_a.log 2>&1; tail -2 gpurun_out/r2_ntt2_a.log
