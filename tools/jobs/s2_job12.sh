set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/s2_pytest_h.log 2>&1; tail -4 gpurun_out/s2_pytest_h.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g2d.json 2> gpurun_out/s2_bench_g2d.err; tail -5 gpurun_out/s2_bench_g2d.err; python - <<'PY'
import json
for l in open('gpurun_out/s2_bench_g2d.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['stage_ms'], d['comm'])
PY
