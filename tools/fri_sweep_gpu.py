"""BASELINE configs[4]: FRI-only sweep (folding 4, remainder max degree 31, Blake3_256, base field) at N GPUs.
The path shards by independent objects here: every rank folds its own codeword (weak scaling, no data-path
collective); times are CUDA events, max over ranks; fold-layer GB/s uses the algorithmic bytes of SURVEY.md 8d
(e*L*1.25 + 32*L per layer, x 4/3 over the layers).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/fri_sweep_gpu.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import winterfell_b200 as wf

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = wf.Context(lr)
rng = np.random.default_rng(100 + rank)
out = []
for log_len in (20, 22, 24, 26):
    L, d, b = 1 << log_len, 1, 8
    poly = rng.integers(0, wf.P, size=(d, L // b), dtype=np.uint64)
    m = ctx.mat_from_host_columns(poly)
    cw = m.lde(3)
    for _ in range(2):  # warm the pool and the twiddle cache
        f, _ = ctx.fri_build_layers_default(wf.HASH_BLAKE3_256, cw, d, 4, 31, b)
        f.free()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        f, roots = ctx.fri_build_layers_default(wf.HASH_BLAKE3_256, cw, d, 4, 31, b)
        f.free()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    alg = (8 * d * L * 1.25 + 32 * L) * 4 / 3
    out.append({"n_gpus": world, "log_len": log_len, "ms_commit_phase_max_over_ranks": round(float(ms[0]), 3),
                "aggregate_fold_GBps": round(world * alg / 1e9 / (float(ms[0]) * 1e-3), 1),
                "frac_of_hbm_per_gpu": round(alg / 1e9 / (float(ms[0]) * 1e-3) / 6571.2, 4)})
    for o in (m, cw):
        o.free()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fri_sweep_r1.jsonl", "a") as fh:
        for r in out:
            print(json.dumps(r))
            fh.write(json.dumps(r) + "\n")
if world > 1:
    dist.destroy_process_group()
