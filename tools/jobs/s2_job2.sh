set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ntt2_pass" -c 6 -o gpurun_out/s2_ntt2_a python tools/prof_prove.py 22 4 1 dev 3 > gpurun_out/s2_ntt2_a.log 2>&1; tail -2 gpurun_out/s2_ntt2_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hash_rows_blake3_kernel|deep_sum|deep_div|fib_constraints|ood_partial|ood_reduce|pow_table" -c 8 -o gpurun_out/s2_cubic_a python tools/prof_prove.py 20 32 1 dev 3 > gpurun_out/s2_cubic_a.log 2>&1; tail -2 gpurun_out/s2_cubic_a.log
./winterfell_b200/_build/ubench_int > gpurun_out/s2_ubench_int.jsonl 2>&1; tail -20 gpurun_out/s2_ubench_int.jsonl
WF_REF_BUDGET_S=100 timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_ref_a.json 2> gpurun_out/s2_ref_a.err; cat gpurun_out/s2_ref_a.json | head -c 600
