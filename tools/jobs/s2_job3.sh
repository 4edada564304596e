set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -k "partition or extension_field or ntt or lde" > gpurun_out/s2_pytest_c.log 2>&1; tail -5 gpurun_out/s2_pytest_c.log
python tools/bench_ntt.py 22:16 20:8 18:64 > gpurun_out/s2_ntt_base.jsonl 2>&1; cat gpurun_out/s2_ntt_base.jsonl
WF_LIB_PATH=$PWD/winterfell_b200/_var/t512/lib.so python tools/bench_ntt.py 22:16 20:8 18:64 > gpurun_out/s2_ntt_t512.jsonl 2>&1; cat gpurun_out/s2_ntt_t512.jsonl
