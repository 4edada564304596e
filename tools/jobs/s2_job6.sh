set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_g2.json 2> gpurun_out/s2_bench_g2.err; tail -c 3000 gpurun_out/s2_bench_g2.json; tail -5 gpurun_out/s2_bench_g2.err
