mkdir -p gpurun_out
export WF_LIB_PATH=$PWD/winterfell_b200/_var/sq/lib.so
timeout 100 python tools/bench_hash.py 21 8 Rp64_256 3 > gpurun_out/s2_hash_sq.jsonl 2> gpurun_out/s2_hash_sq.err
timeout 60 python tools/bench_hash.py 21 8 RpJive64_256 3 >> gpurun_out/s2_hash_sq.jsonl 2>> gpurun_out/s2_hash_sq.err
cat gpurun_out/s2_hash_sq.jsonl
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prover.py -x -q -m gpu -k "row_hash or partitioned_row or proof_bytes or grinding" > gpurun_out/s2_pytest_sq.log 2>&1; tail -2 gpurun_out/s2_pytest_sq.log
