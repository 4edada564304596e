"""Row hashing + Merkle tree throughput of every hasher (Merkle leaves/s; for the algebraic hashers
permutations/s against the integer-pipe ceiling of their S-box arithmetic). CUDA events on the context stream, L2 flushed
between repetitions.
   python tools/bench_hash.py [log_rows] [cols] [hasher name, e.g. Rp64_256: only that one] [repetitions]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import winterfell_b200 as wf

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 21
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 8
only = sys.argv[3] if len(sys.argv) > 3 else None
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rows = 1 << log_rows
stream = torch.cuda.Stream()
ctx = wf.Context(0, stream.cuda_stream)
rng = np.random.default_rng(3)
NAMES = {wf.HASH_BLAKE3_256: "Blake3_256", wf.HASH_RP64_256: "Rp64_256", wf.HASH_RPJIVE64_256: "RpJive64_256",
         wf.HASH_BLAKE3_192: "Blake3_192", wf.HASH_SHA3_256: "Sha3_256"}
# permutations (compressions) per row hash and per tree node
def calls(h, cols):
    if h in (wf.HASH_BLAKE3_256, wf.HASH_BLAKE3_192):
        return max(1, -(-cols * 8 // 64)), 1
    if h == wf.HASH_RP64_256:
        return max(1, -(-cols // 8)), 1
    if h == wf.HASH_RPJIVE64_256:
        return max(1, -(-cols // 4)), 1
    return max(1, -(-(cols * 8 + 1) // 136)), 1     # Sha3_256: 136-byte rate blocks
with torch.cuda.stream(stream):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    m = ctx.mat_from_host_columns(rng.integers(0, wf.P, size=(cols, rows), dtype=np.uint64))
    for h in (wf.HASH_BLAKE3_256, wf.HASH_BLAKE3_192, wf.HASH_SHA3_256, wf.HASH_RP64_256, wf.HASH_RPJIVE64_256):
        if only and NAMES[h] != only:
            continue
        times = []
        for rep in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            t = ctx.commit_rows(h, m)
            b.record(stream)
            b.synchronize()
            times.append(a.elapsed_time(b))
            root = bytes(t.root())
            t.free()
        ms = min(times[1:]) if len(times) > 1 else times[0]
        per_row, per_node = calls(h, cols)
        n_calls = rows * per_row + (rows - 1) * per_node
        print(json.dumps({"hasher": NAMES[h], "rows": rows, "cols": cols, "commit_ms": round(ms, 4),
                          "leaves_per_s": round(rows / (ms * 1e-3), 1), "permutations_or_compressions": n_calls,
                          "G_per_s": round(n_calls / (ms * 1e-3) / 1e9, 4), "root": root.hex()[:16]}), flush=True)
    m.free()
