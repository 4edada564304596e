set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_j.log 2>&1; tail -5 gpurun_out/s2_pytest_j.log
for v in default c1 c4 ds128 ds512; do
  if [ $v = default ]; then unset WF_LIB_PATH; else export WF_LIB_PATH=$PWD/winterfell_b200/_var/$v/lib.so; fi
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_var_$v.json 2> gpurun_out/s2_bench_var_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/s2_bench_var_$v.json').read().strip().splitlines()[-1])
print('$v', d['value'], d['e2e']['value'], d['stage_ms'])
PY
done
unset WF_LIB_PATH
