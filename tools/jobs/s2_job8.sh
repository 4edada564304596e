set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g8.json 2> gpurun_out/s2_bench_g8.err; tail -c 2500 gpurun_out/s2_bench_g8.json; tail -5 gpurun_out/s2_bench_g8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-record > gpurun_out/s2_bench_g4.json 2> gpurun_out/s2_bench_g4.err; tail -c 2500 gpurun_out/s2_bench_g4.json; tail -5 gpurun_out/s2_bench_g4.err
