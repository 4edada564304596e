set -x
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/memory.max; free -g | head -2; grep -m1 "model name" /proc/cpuinfo; nvidia-smi -L | head -3
./winterfell_b200/_build/ubench_int > gpurun_out/r2_ubench_int.jsonl 2>&1
WF_STAGES=1 timeout 600 python tools/prof_prove.py 22 32 3 dev 3 > gpurun_out/r2_base_cfg3.log 2>&1
tail -3 gpurun_out/r2_base_cfg3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"deep_|ood_|fib_constraints|hash_rows|merkle|fri_" -c 24 -o gpurun_out/r2_base_cubic python tools/prof_prove.py 20 32 1 dev 3 > gpurun_out/r2_base_ncu.log 2>&1
tail -2 gpurun_out/r2_base_ncu.log
