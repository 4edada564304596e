set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_e.log 2>&1; tail -5 gpurun_out/s2_pytest_e.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_d.json 2> gpurun_out/s2_bench_d.err; tail -3 gpurun_out/s2_bench_d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_d.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms'])
print(d['cfg2']['value'], d['cfg2']['e2e'], d['cfg2']['stage_ms'])
print(d.get('fri_sweep'))
PY
