set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_prover.py -x -q > gpurun_out/s2_pytest_i.log 2>&1; tail -4 gpurun_out/s2_pytest_i.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g2e.json 2> gpurun_out/s2_bench_g2e.err; tail -5 gpurun_out/s2_bench_g2e.err; python - <<'PY'
import json
for l in open('gpurun_out/s2_bench_g2e.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['stage_ms'], d['comm'])
PY
timeout 600 python tools/jit_bench.py 16 > gpurun_out/s2_jit_bench.jsonl 2>&1; cat gpurun_out/s2_jit_bench.jsonl
