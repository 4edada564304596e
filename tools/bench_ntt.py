"""Kernel experiment driver: times wf_mat_interpolate + wf_mat_lde (the NTT pass kernels alone) with CUDA events on the context stream
for a list of shapes and prints one JSON line per shape with a checksum of the LDE (so that variant builds selected through
WF_LIB_PATH can be compared for speed AND equality).   python tools/bench_ntt.py [log_n:cols ...]"""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import winterfell_b200 as wf

shapes = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(22, 16), (20, 8)]
stream = torch.cuda.Stream()
ctx = wf.Context(0, stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for log_n, cols in shapes:
    n = 1 << log_n
    rng = np.random.default_rng(log_n * 100 + cols)
    tr = rng.integers(0, wf.P, size=(cols, n), dtype=np.uint64)
    with torch.cuda.stream(stream):
        m = ctx.mat_from_host_columns(tr)
        ti, tl = [], []
        for rep in range(6):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            polys = m.interpolate()
            e[1].record(stream)
            lde = polys.lde(3)
            e[2].record(stream)
            e[2].synchronize()
            if rep >= 2:
                ti.append(e[0].elapsed_time(e[1])); tl.append(e[1].elapsed_time(e[2]))
            if rep == 5:
                # checksum over a strided sample of LDE rows (reading the whole LDE back would dominate the run)
                idx = np.arange(0, n * 8, max(1, (n * 8) // 4096), dtype=np.uint64)
                h = hashlib.sha256(lde.read_rows(idx).tobytes() + polys.read_rows(np.arange(0, n, max(1, n // 1024), dtype=np.uint64)).tobytes()).hexdigest()[:16]
            polys.free(); lde.free()
        m.free()
    N = n * 8
    print(json.dumps({"lib": os.environ.get("WF_LIB_PATH", "default"), "log_n": log_n, "cols": cols, "interp_ms": round(min(ti), 4),
                      "lde_ms": round(min(tl), 4), "lde_Gelem_s": round(N * cols / min(tl) / 1e6, 2),
                      "ntt_Gelem_s": round((N + n) * cols / (min(tl) + min(ti)) / 1e6, 2), "checksum": h}), flush=True)
ctx.close()
