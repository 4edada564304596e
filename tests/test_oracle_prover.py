"""CPU tests of the oracle's generate_proof / verify restatement (oracle/wf_prover.cpp): round trips,
negative tests, and the reference's own end-to-end test configuration
(examples/src/fibonacci/fib_small/tests.rs:8-24, examples/src/tests.rs:8-17)."""
import numpy as np
import pytest

import airs


@pytest.mark.parametrize("ext", [1, 2])
def test_reference_fib_small_test_config(oracle, ext):
    # FibSmall, sequence length 128 => 64 rows? the reference example proves `sequence_length / 2` rows:
    # fib_small::get_example(&options, 128) builds a 64-row trace; options: 28 queries, blowup 8,
    # grinding 0, folding 4, remainder max degree 7, Rp64_256 (fibonacci/utils.rs:33-42, fib_small/tests.rs)
    for n in (64, 128):
        trace, res = oracle.build_fib_trace(1, n)
        opts = oracle.make_opts(num_queries=28, blowup=8, grinding=0, ext=ext, folding=4, rem_max_deg=7, hash_id=oracle.RP64)
        proof = oracle.prove_fib(trace, res, opts)
        assert oracle.verify_fib(proof, 1, res, oracle.RP64) == 0
        # wrong public input must be rejected (examples/src/tests.rs:8-17)
        assert oracle.verify_fib(proof, 1, res + np.uint64(1), oracle.RP64) != 0


@pytest.mark.parametrize("k,n,ext,h,fold,batch", [(1, 256, 1, 0, 8, 0), (4, 256, 3, 0, 4, 1), (2, 64, 3, 0, 2, 2), (3, 128, 2, 1, 16, 0)])
def test_fib_round_trip_and_tamper(oracle, k, n, ext, h, fold, batch):
    trace, res = oracle.build_fib_trace(k, n)
    opts = oracle.make_opts(ext=ext, hash_id=h, folding=fold, batch_c=batch, batch_d=batch, grinding=4)
    proof = oracle.prove_fib(trace, res, opts)
    assert oracle.verify_fib(proof, k, res, h) == 0
    rng = np.random.default_rng(k * n)
    for _ in range(12):                                    # any flipped bit must be caught
        t = bytearray(proof)
        i = int(rng.integers(0, len(t)))
        t[i] ^= 1 << int(rng.integers(0, 8))
        assert oracle.verify_fib(bytes(t), k, res, h) != 0, i
    assert oracle.verify_fib(proof[:-1], k, res, h) != 0   # truncated
    assert oracle.verify_fib(proof + b"\0", k, res, h) != 0  # trailing byte


def test_generic_air_equals_specialised_fib(oracle):
    for k, n, ext in [(1, 64, 1), (4, 128, 3)]:
        desc, trace = airs.fib_small_x(k, n)
        tr2, res = oracle.build_fib_trace(k, n)
        assert (trace == tr2).all()
        opts = oracle.make_opts(ext=ext, grinding=2)
        assert oracle.prove_air(desc, trace, opts) == oracle.prove_fib(tr2, res, opts)


@pytest.mark.parametrize("name", ["mulfib2", "periodic_mix"])
@pytest.mark.parametrize("ext", [1, 2, 3])
def test_generic_airs_round_trip(oracle, name, ext):
    desc, trace = getattr(airs, name)(256)
    opts = oracle.make_opts(ext=ext, grinding=2, blowup=8, folding=4, rem_max_deg=7)
    proof = oracle.prove_air(desc, trace, opts)
    assert oracle.verify_air(desc, proof) == 0
    bad = desc.copy()
    bad[-2] ^= np.uint64(1)
    assert oracle.verify_air(bad, proof) != 0
    # an invalid trace (one flipped cell) must not verify: the constraint quotient stops being low degree
    t2 = trace.copy()
    t2[0, 17] ^= np.uint64(1)
    assert oracle.verify_air(desc, oracle.prove_air(desc, t2, opts)) != 0


# ---- auxiliary trace segment ----
@pytest.mark.parametrize("ext", [1, 2, 3])
def test_aux_segment_roundtrip_and_tamper(oracle, ext):
    desc, trace, builder = airs.perm_rap(64)
    opts = oracle.make_opts(num_queries=16, blowup=8, grinding=2, ext=ext, folding=4, rem_max_deg=7)
    proof = oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert oracle.verify_air(desc, proof) == 0
    # Context carries (main width, aux width, aux rands): air/src/air/trace_info.rs:240-264
    assert proof[0] == 3 and proof[1] == airs.PERM_RAP_AUX_WIDTH and proof[2] == 2
    b = bytearray(proof)
    b[len(b) // 2] ^= 1
    assert oracle.verify_air(desc, bytes(b)) != 0
    # main trace inconsistent with the aux columns: the permutation argument fails at the OOD check
    bad = trace.copy()
    bad[2, 3] = (int(bad[2, 3]) + 1) % airs.P
    assert oracle.verify_air(desc, oracle.prove_air_aux(desc, bad, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)) != 0
    # a single-segment verifier description does not accept a two-segment proof
    single, _ = airs.fib_small_x(1, 64)
    assert oracle.verify_air(single, proof) != 0


# ---- sequence assertions ----
@pytest.mark.parametrize("n", [32, 256])
def test_sequence_assertions_roundtrip(oracle, n):
    desc, trace = airs.sequence_mix(n)
    opts = oracle.make_opts(num_queries=16, blowup=8, ext=2, folding=4, rem_max_deg=7)
    proof = oracle.prove_air(desc, trace, opts)
    assert oracle.verify_air(desc, proof) == 0
    bad = desc.copy()
    i = int(np.nonzero(bad == np.uint64(trace[0, 5]))[0][0])
    bad[i] ^= np.uint64(1)
    assert oracle.verify_air(bad, proof) != 0
    # the prover run on a trace that violates an asserted sequence value yields a rejected proof
    t2 = trace.copy()
    t2[2, 1] = 8  # periodic assertion (column 2, first step 1) expects 7; column 2 has no transition constraint
    assert oracle.verify_air(desc, oracle.prove_air(desc, t2, opts)) != 0


# ---- committed golden fixtures (tests/golden/proof_digests.json, made by tests/golden/make_proof_digests.py) ----
def _golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof_digests.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("idx", range(11))
def test_oracle_reproduces_golden_proof_digests(oracle, idx):
    import hashlib
    rec = _golden()[idx]
    opts = oracle.make_opts(**rec["opts"])
    if rec["kind"] == "fib":
        trace, res = oracle.build_fib_trace(rec["k"], 1 << rec["log_n"])
        proof = oracle.prove_fib(trace, res, opts)
    elif rec["air"] == "perm_rap":
        desc, trace, builder = airs.perm_rap(rec["n"])
        proof = oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    else:
        desc, trace = getattr(airs, rec["air"])(rec["n"])
        proof = oracle.prove_air(desc, trace, opts)
    assert len(proof) == rec["bytes"] and hashlib.sha256(proof).hexdigest() == rec["sha256"]


@pytest.mark.parametrize("k,ext,h,parts,rate", [(4, 1, 0, 2, 8), (4, 3, 0, 4, 8), (8, 1, 0, 8, 8), (4, 2, 1, 2, 8), (1, 3, 0, 4, 64)])
def test_partitioned_commitments_round_trip(oracle, k, ext, h, parts, rate):
    # ProofOptions::with_partitions (air/src/options.rs:193-200): rows are hashed as merge_many of the digests of their
    # column partitions (row_matrix.rs:204-223) and the verifier re-hashes the queried rows the same way
    # (verifier/src/channel.rs:431-453); the last case has partition_size > row width: ONE partition, still merge_many
    n = 256
    trace, res = oracle.build_fib_trace(k, n)
    plain = oracle.prove_fib(trace, res, oracle.make_opts(ext=ext, hash_id=h, grinding=2))
    opts = oracle.make_opts(ext=ext, hash_id=h, grinding=2, num_partitions=parts, hash_rate=rate)
    proof = oracle.prove_fib(trace, res, opts)
    assert oracle.verify_fib(proof, k, res, h) == 0
    assert proof != plain
    # the partition options travel in the proof's serialized ProofOptions (options.rs:318-319): a verifier told otherwise
    # re-hashes the rows differently and must reject
    i = next(j for j in range(len(proof)) if proof[j] != plain[j])
    assert (proof[i], proof[i + 1]) == (parts, rate) and (plain[i], plain[i + 1]) == (1, 1)
    t = bytearray(proof)
    t[i], t[i + 1] = 1, 1
    assert oracle.verify_fib(bytes(t), k, res, h) != 0


@pytest.mark.parametrize("ext", [1, 3])
def test_aux_assertions_depending_on_random_elements(oracle, ext):
    # Air::get_aux_assertions receives the AuxRandElements (air/src/air/mod.rs:279): an assertion value that is a function of
    # the drawn elements is supplied by a callback at the same transcript point in the prover and in the verifier
    n = 64
    desc, trace, builder = airs.perm_rap(n, dyn_last_q=True)
    opts = oracle.make_opts(num_queries=16, grinding=2, ext=ext, folding=4, rem_max_deg=7)
    nv = builder.num_values
    proof = oracle.prove_air_aux_dyn(desc, trace, opts, builder, builder.values_fn, airs.PERM_RAP_AUX_WIDTH, 2, nv)
    assert oracle.verify_air_dyn(desc, proof, 0, builder.values_fn, 2, nv, ext) == 0
    # the placeholder value of the description (0) is not what the trace satisfies: a verifier without the callback rejects,
    # and so does one whose callback computes a different value
    assert oracle.verify_air(desc, proof, 0) != 0
    wrong = lambda rand, values: np.where(np.arange(nv)[:, None] == 3, values + np.uint64(1), builder.values_fn(rand, values))
    assert oracle.verify_air_dyn(desc, proof, 0, wrong, 2, nv, ext) != 0
