set -x
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max; free -g | head -2; nvidia-smi -L | head -2
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_a.log 2>&1; tail -5 gpurun_out/s2_pytest_a.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s2_bench_a.json 2> gpurun_out/s2_bench_a.err; tail -c 600 gpurun_out/s2_bench_a.json; tail -5 gpurun_out/s2_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_cfg3.csv python tools/prof_prove.py 22 32 1 dev 3 > gpurun_out/s2_launches_cfg3.log 2>&1; tail -2 gpurun_out/s2_launches_cfg3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_cfg2.csv python tools/prof_prove.py 20 4 1 dev 1 > gpurun_out/s2_launches_cfg2.log 2>&1; tail -2 gpurun_out/s2_launches_cfg2.log
