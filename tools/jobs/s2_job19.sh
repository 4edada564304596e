set -x
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:alg_kernel -c 3 -f -o gpurun_out/r2_rp64_kernels python tools/bench_hash.py 20 8 Rp64_256 1 > gpurun_out/r2_rp64_ncu.log 2>&1; tail -3 gpurun_out/r2_rp64_ncu.log
ls -la gpurun_out/r2_rp64_kernels.ncu-rep
