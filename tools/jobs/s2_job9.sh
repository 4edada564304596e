set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s2_pytest_f.log 2>&1; tail -5 gpurun_out/s2_pytest_f.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g2b.json 2> gpurun_out/s2_bench_g2b.err; tail -5 gpurun_out/s2_bench_g2b.err; python - <<'PY'
import json
for l in open('gpurun_out/s2_bench_g2b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['stage_ms'], d['comm'])
PY
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_e.json 2> gpurun_out/s2_bench_e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_bench_e.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms'])
PY
