set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python tools/bench_hash.py 21 8 > gpurun_out/s2_hash_new.jsonl 2> gpurun_out/s2_hash_new.err; cat gpurun_out/s2_hash_new.jsonl; tail -2 gpurun_out/s2_hash_new.err
WF_LIB_PATH=$PWD/winterfell_b200/_var/rp_old/lib.so timeout 300 python tools/bench_hash.py 21 8 > gpurun_out/s2_hash_old.jsonl 2> gpurun_out/s2_hash_old.err; cat gpurun_out/s2_hash_old.jsonl; tail -2 gpurun_out/s2_hash_old.err
timeout 300 python tools/bench_hash.py 21 64 > gpurun_out/s2_hash_new64.jsonl 2>&1; cat gpurun_out/s2_hash_new64.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_j17.json 2> gpurun_out/s2_bench_j17.err; tail -c 600 gpurun_out/s2_bench_j17.json
