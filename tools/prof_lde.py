"""Short driver for ncu captures: one trace commit (interpolate + LDE + hash + Merkle) at 2^log_n x cols."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import winterfell_b200 as wf
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = wf.Context(0)
rng = np.random.default_rng(1)
tr = rng.integers(0, wf.P, size=(cols, 1 << log_n), dtype=np.uint64)
m = ctx.mat_from_host_columns(tr)
for _ in range(reps):
    polys = m.interpolate()
    lde = polys.lde(3)
    tree = ctx.commit_rows(wf.HASH_BLAKE3_256, lde)
    fm = lde.select_columns(0, 1)
    f, roots = ctx.fri_build_layers_default(wf.HASH_BLAKE3_256, fm, 1, 4, 31, 8)
    for h in (polys, lde, tree, fm, f):
        h.free()
ctx.sync()
print("ok", tree)
