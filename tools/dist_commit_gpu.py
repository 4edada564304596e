"""N-GPU check + timing of the column-sharded trace commitment (run under torchrun on a GPU box):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_commit_gpu.py [log_n] [cols]
Rank 0 compares the distributed root with the single-GPU commitment of the whole trace."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import winterfell_b200 as wf
from winterfell_b200 import dist as wd

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = wf.Context(lr)
n = 1 << log_n
rng = np.random.default_rng(5)
trace = rng.integers(0, wf.P, size=(cols, n), dtype=np.uint64)
lo, hi = wd.column_range(cols, world, rank)
local = torch.from_numpy(trace[lo:hi].copy().view(np.int64)).cuda()
be = wd.CudaBackend(ctx)
root, rows, dg, nd = wd.sharded_trace_commit(be, wf.HASH_BLAKE3_256, local, cols, log_n, 3)   # warm-up
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    root, rows, dg, nd = wd.sharded_trace_commit(be, wf.HASH_BLAKE3_256, local, cols, log_n, 3)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    import json
    m = ctx.mat_from_host_columns(trace)
    def single():
        p_ = m.interpolate(); l_ = p_.lde(3); t_ = ctx.commit_rows(wf.HASH_BLAKE3_256, l_)
        r_ = t_.root()
        for o in (p_, l_, t_):
            o.free()
        return r_
    ref_root = single()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(3):
        single()
    s1.record(); torch.cuda.synchronize()
    N = n << 3
    rec = {"world": world, "log_n": log_n, "cols": cols, "blowup": 8, "hash": "blake3_256",
           "root_matches_single_gpu": ref_root == root, "ms_sharded_max_over_ranks": round(float(ms[0]), 3),
           "ms_single_gpu_same_box": round(s0.elapsed_time(s1) / 3, 3),
           "lde_gelem_per_s_sharded": round(N * cols / (float(ms[0]) * 1e-3) / 1e9, 2),
           "leaves_per_s_sharded": round(N / (float(ms[0]) * 1e-3), 1)}
    print(json.dumps(rec))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sharded_commit_r1.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
dist.destroy_process_group()
