"""Constraint evaluation of the heavy rescue_like AIR (tests/airs.py: degree-7 S-boxes + MDS layer on 6 columns) through the
built-in interpreter and through the kernel NVRTC compiles for the AIR; CUDA events on the context stream.
   python tools/jit_bench.py [log_n]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import winterfell_b200 as wf, airs

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << log_n
desc, trace = airs.rescue_like(n)
stream = torch.cuda.Stream()
ctx = wf.Context(0, stream.cuda_stream)
for ext in (1, 3):
    width = trace.shape[0]
    ncoef = width + width + 1            # transition constraints + assertions (width single + 1)
    rng = np.random.default_rng(5)
    coeffs = rng.integers(0, wf.P, size=(ncoef, ext), dtype=np.uint64)
    with torch.cuda.stream(stream):
        m = ctx.mat_from_host_columns(trace)
        lde = m.interpolate().lde(3)
        res = {}
        for mode in ("jit", "interpreter"):
            ctx.set_jit(mode == "jit")
            times = []
            for rep in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                out = ctx.eval_constraints(desc, log_n, 8, ext, lde, None, coeffs)
                b.record(stream)
                b.synchronize()
                times.append(a.elapsed_time(b))
                if rep == 4:
                    res[mode + "_checksum"] = int(np.bitwise_xor.reduce(out.to_rows().ravel()))
                out.free()
            res[mode + "_ms"] = round(min(times[1:]), 4)
        ctx.set_jit(True)
    res.update({"air": "rescue_like (6 columns, degree 7)", "log_n": log_n, "ext": ext, "ce_rows": n * 8, "jit_stats": ctx.jit_stats(),
                "speedup": round(res["interpreter_ms"] / res["jit_ms"], 2), "equal": res["jit_checksum"] == res["interpreter_checksum"]})
    print(json.dumps(res), flush=True)
ctx.close()
