set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_d.log 2>&1; tail -5 gpurun_out/s2_pytest_d.log
python tools/bench_ntt.py 22:16 20:8 > gpurun_out/s2_ntt_base2.jsonl 2>&1; cat gpurun_out/s2_ntt_base2.jsonl
WF_LIB_PATH=$PWD/winterfell_b200/_var/rteps/lib.so python tools/bench_ntt.py 22:16 20:8 > gpurun_out/s2_ntt_rteps.jsonl 2>&1; cat gpurun_out/s2_ntt_rteps.jsonl
WF_LIB_PATH=$PWD/winterfell_b200/_var/rteps/lib.so timeout 900 python -m pytest tests/ -x -q -m gpu -k "field_ops or extension_field or ntt or lde or proof_bytes" > gpurun_out/s2_pytest_rteps.log 2>&1; tail -3 gpurun_out/s2_pytest_rteps.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_b.json 2> gpurun_out/s2_bench_b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms'])
print(d['cfg2']['value'], d['cfg2']['e2e'], d['cfg2']['stage_ms'])
PY
WF_LIB_PATH=$PWD/winterfell_b200/_var/rteps/lib.so timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_b_rteps.json 2> gpurun_out/s2_bench_b_rteps.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_b_rteps.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms'])
print(d['cfg2']['value'], d['cfg2']['e2e'], d['cfg2']['stage_ms'])
PY
