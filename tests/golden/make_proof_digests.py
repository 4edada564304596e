"""Regenerates tests/golden/proof_digests.json: SHA-256 and length of the serialized proofs the ORACLE prover
(oracle/wf_prover.cpp, the CPU restatement of winterfell v0.13.1 generate_proof) emits for a fixed list of
small configurations. The reference holds no golden proof bytes and cannot be built here (Rust, no cargo),
so these fixtures pin the restatement itself: the CPU suite checks that the oracle still reproduces them,
the GPU suite that the device pipeline does.
    python tests/golden/make_proof_digests.py        (from the repository root)"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import airs  # noqa: E402
from oracle import oracle as O  # noqa: E402

FIB = [  # k, log_n, options
    (1, 7, dict(num_queries=28, blowup=8, grinding=0, ext=1, folding=4, rem_max_deg=7, hash_id=1)),   # examples/src/fibonacci/fib_small/tests.rs
    (1, 7, dict(num_queries=28, blowup=8, grinding=0, ext=2, folding=4, rem_max_deg=7, hash_id=1)),
    (1, 10, dict(num_queries=28, blowup=8, grinding=8, ext=1, folding=8, rem_max_deg=31, hash_id=0)),
    (4, 10, dict(num_queries=32, blowup=8, grinding=8, ext=1, folding=4, rem_max_deg=31, hash_id=0)),
    (4, 9, dict(num_queries=32, blowup=8, grinding=4, ext=3, folding=4, rem_max_deg=31, batch_c=1, batch_d=1, hash_id=0)),
    (32, 8, dict(num_queries=32, blowup=8, grinding=0, ext=3, folding=4, rem_max_deg=31, batch_c=2, batch_d=2, hash_id=0)),
    (1, 3, dict(num_queries=5, blowup=128, grinding=0, ext=1, folding=16, rem_max_deg=7, hash_id=0)),
]
AIRS = [  # builder name, n, options
    ("mulfib2", 256, dict(num_queries=24, blowup=8, grinding=2, ext=1, folding=4, rem_max_deg=15, hash_id=0)),
    ("periodic_mix", 512, dict(num_queries=24, blowup=8, grinding=2, ext=3, folding=4, rem_max_deg=15, batch_c=1, batch_d=1, hash_id=0)),
    ("sequence_mix", 256, dict(num_queries=20, blowup=8, grinding=1, ext=2, folding=4, rem_max_deg=7, hash_id=0)),
    ("perm_rap", 128, dict(num_queries=20, blowup=8, grinding=2, ext=2, folding=4, rem_max_deg=7, batch_c=2, batch_d=2, hash_id=1)),
]


def entries():
    for k, log_n, kw in FIB:
        trace, res = O.build_fib_trace(k, 1 << log_n)
        yield {"kind": "fib", "k": k, "log_n": log_n, "opts": kw}, O.prove_fib(trace, res, O.make_opts(**kw))
    for name, n, kw in AIRS:
        out = getattr(airs, name)(n)
        opts = O.make_opts(**kw)
        if name == "perm_rap":
            desc, trace, builder = out
            proof = O.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
        else:
            desc, trace = out
            proof = O.prove_air(desc, trace, opts)
        yield {"kind": "air", "air": name, "n": n, "opts": kw}, proof


if __name__ == "__main__":
    recs = []
    for cfg, proof in entries():
        cfg.update(bytes=len(proof), sha256=hashlib.sha256(proof).hexdigest())
        recs.append(cfg)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "proof_digests.json"), "w") as f:
        json.dump(recs, f, indent=1)
    print(f"wrote {len(recs)} digests")
