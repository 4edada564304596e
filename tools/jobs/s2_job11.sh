set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g2c.json 2> gpurun_out/s2_bench_g2c.err; tail -5 gpurun_out/s2_bench_g2c.err; python - <<'PY'
import json
for l in open('gpurun_out/s2_bench_g2c.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['stage_ms'], d['comm'])
PY
timeout 300 python -m pytest tests/test_gpu_prover.py -x -q -k "depending_on_random" 2>&1 | tail -3
