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
    m = ctx.mat_from_host_columns(trace)
    t = ctx.commit_rows(wf.HASH_BLAKE3_256, m.interpolate().lde(3))
    print({"world": world, "log_n": log_n, "cols": cols, "root_matches_single_gpu": t.root() == root, "ms_max_over_ranks": round(float(ms[0]), 3)})
dist.destroy_process_group()
