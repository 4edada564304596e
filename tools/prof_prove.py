"""Short driver for ncu captures: full proofs of FibSmall x pairs at 2^log_n rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import winterfell_b200 as wf
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
resident = len(sys.argv) > 4 and sys.argv[4] == "dev"   # trace already in HBM (bench.py's `value` arm)
P = wf.P
n = 1 << log_n
trace = np.zeros((2 * pairs, n), dtype=np.uint64)
res = np.zeros(pairs, dtype=np.uint64)
for j in range(pairs):
    va = vb = j + 1
    ca, cb = [0] * n, [0] * n
    for i in range(n):
        ca[i], cb[i] = va, vb
        va = (va + vb) % P
        vb = (vb + va) % P
    trace[2 * j], trace[2 * j + 1], res[j] = np.array(ca, dtype=np.uint64), np.array(cb, dtype=np.uint64), cb[n - 1]
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
