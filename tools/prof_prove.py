"""Short driver for ncu captures: full proofs of FibSmall x pairs at 2^log_n rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import winterfell_b200 as wf
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
resident = len(sys.argv) > 4 and sys.argv[4] == "dev"   # trace already in HBM (bench.py's `value` arm)
n = 1 << log_n
trace, res = wf.build_fib_trace(pairs, n)
ctx = wf.Context(0)
ext = int(sys.argv[5]) if len(sys.argv) > 5 else 1
opts = np.array([32, 8, 16, ext, 4, 31, 0, 0, 0], dtype=np.uint32)
if resident:
    import torch
    d_trace = torch.from_numpy(trace.view(np.int64)).cuda()
for _ in range(reps):
    proof = ctx.prove_fib_dev(d_trace.data_ptr(), pairs, log_n, res, opts) if resident else ctx.prove_fib(trace, res, opts)
if os.environ.get("WF_STAGES"):
    ctx.set_profiling(True)
    proof = ctx.prove_fib_dev(d_trace.data_ptr(), pairs, log_n, res, opts) if resident else ctx.prove_fib(trace, res, opts)
    print({k: round(v, 3) for k, v in ctx.stage_times()})
print("ok", len(proof))
