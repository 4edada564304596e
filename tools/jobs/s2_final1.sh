set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_line_1gpu.json 2> gpurun_out/r2_bench_line_1gpu.err; tail -c 400 gpurun_out/r2_bench_line_1gpu.json; tail -3 gpurun_out/r2_bench_line_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_reference_line.json 2> gpurun_out/r2_reference_line.err; head -c 700 gpurun_out/r2_reference_line.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3_final.csv python tools/prof_prove.py 22 32 1 dev 3 > gpurun_out/r2_launches_cfg3_final.log 2>&1; tail -1 gpurun_out/r2_launches_cfg3_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
