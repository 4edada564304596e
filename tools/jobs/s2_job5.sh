set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^deep_sum|^deep_div|^fib_constraints|^ood_partial|^ood_reduce|^cols_to_seg|^fri_hash|^fri_fold" -c 9 -o gpurun_out/s2_cubic_b python tools/prof_prove.py 20 32 1 dev 3 > gpurun_out/s2_cubic_b.log 2>&1; tail -2 gpurun_out/s2_cubic_b.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"ntt2_pass" --csv --log-file gpurun_out/s2_ntt_traffic_cfg3.csv python tools/prof_prove.py 22 32 1 dev 3 > gpurun_out/s2_ntt_traffic_cfg3.log 2>&1; tail -1 gpurun_out/s2_ntt_traffic_cfg3.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"ntt2_pass" --csv --log-file gpurun_out/s2_ntt_traffic_cfg2.csv python tools/prof_prove.py 20 4 1 dev 1 > gpurun_out/s2_ntt_traffic_cfg2.log 2>&1; tail -1 gpurun_out/s2_ntt_traffic_cfg2.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_c.json 2> gpurun_out/s2_bench_c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms'])
print(d['cfg2']['value'], d['cfg2']['e2e'], d['cfg2']['stage_ms'])
PY
