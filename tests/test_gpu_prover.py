"""GPU proof generation vs the oracle: the serialized Proof produced by the device pipeline
(wf_prove_fib) must be byte-identical to the CPU restatement of Prover::generate_proof, and the
restated verifier (verifier/src/lib.rs) must accept it. Mirrors examples/src/fibonacci/fib_small/tests.rs:8-24
(n = 128, Rp64_256, base and quadratic extension, 28 queries, blowup 8, folding 4, remainder 7) and
examples/src/tests.rs:8-17 (wrong public input is rejected)."""
import numpy as np
import pytest

import winterfell_b200 as wf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = wf.Context(0)
    yield c
    c.close()


CASES = [
    # k, log_n, ext, hash, folding, rem_max_deg, batching, grinding, blowup, queries
    (1, 7, 1, wf.HASH_RP64_256, 4, 7, 0, 0, 8, 28),      # the reference's own fib_small test
    (1, 7, 2, wf.HASH_RP64_256, 4, 7, 0, 0, 8, 28),      # ... with the quadratic extension
    (1, 10, 1, wf.HASH_BLAKE3_256, 8, 31, 0, 8, 8, 28),  # configs[0] shape (fib_small CLI defaults), small n
    (1, 8, 3, wf.HASH_BLAKE3_256, 4, 7, 0, 4, 8, 28),
    (4, 10, 1, wf.HASH_BLAKE3_256, 4, 31, 0, 8, 8, 32),  # configs[1] AIR (8 columns), small n
    (4, 12, 3, wf.HASH_BLAKE3_256, 4, 31, 1, 4, 8, 32),  # two-pass NTT sizes, cubic, algebraic batching
    (32, 9, 3, wf.HASH_BLAKE3_256, 4, 31, 2, 0, 8, 32),  # configs[2] AIR (64 columns, cubic), Horner batching
    (2, 8, 2, wf.HASH_BLAKE3_256, 2, 7, 0, 0, 4, 20),
    (3, 9, 1, wf.HASH_RP64_256, 16, 7, 2, 0, 16, 12),
    (1, 7, 1, wf.HASH_RPJIVE64_256, 4, 7, 0, 0, 8, 28),  # fib_small with the Jive-mode hasher (rp64_256_jive/mod.rs)
    (2, 9, 3, wf.HASH_RPJIVE64_256, 4, 7, 1, 3, 8, 20),
    (4, 8, 2, wf.HASH_RPJIVE64_256, 8, 15, 2, 0, 4, 16),
    (1, 7, 1, wf.HASH_BLAKE3_192, 4, 7, 0, 0, 8, 28),    # 24-byte digests (crypto/src/hash/blake/mod.rs:73-123)
    (4, 10, 3, wf.HASH_BLAKE3_192, 4, 31, 1, 6, 8, 32),
    (32, 9, 2, wf.HASH_BLAKE3_192, 8, 7, 2, 0, 8, 20),
    (1, 7, 1, wf.HASH_SHA3_256, 4, 7, 0, 0, 8, 28),      # Sha3_256 (crypto/src/hash/sha/mod.rs:19)
    (4, 10, 3, wf.HASH_SHA3_256, 4, 31, 1, 5, 8, 32),
    (16, 9, 2, wf.HASH_SHA3_256, 8, 7, 2, 0, 8, 20),     # 32 columns: rows longer than one 136-byte rate block
]


@pytest.mark.parametrize("k,log_n,ext,h,fold,rem,batch,grind,blowup,nq", CASES)
def test_proof_bytes_match_oracle_and_verify(ctx, oracle, k, log_n, ext, h, fold, rem, batch, grind, blowup, nq):
    n = 1 << log_n
    trace, results = oracle.build_fib_trace(k, n)
    opts = oracle.make_opts(num_queries=nq, blowup=blowup, grinding=grind, ext=ext, folding=fold, rem_max_deg=rem,
                            batch_c=batch, batch_d=batch, hash_id=h)
    want = oracle.prove_fib(trace, results, opts)
    got = ctx.prove_fib(trace, results, opts)
    assert len(got) == len(want)
    assert got == want, next(i for i in range(len(got)) if got[i] != want[i])
    assert oracle.verify_fib(got, k, results, h) == 0
    # negative tests: wrong public input (examples/src/tests.rs:8-17) and a flipped byte
    bad = results.copy()
    bad[0] += np.uint64(1)
    assert oracle.verify_fib(got, k, bad, h) != 0
    t = bytearray(got)
    t[len(t) // 3] ^= 0x40
    assert oracle.verify_fib(bytes(t), k, results, h) != 0


def test_montgomery_trace_input(ctx, oracle):
    # the Rust shim passes &[BaseElement] reinterpreted as u64: Montgomery words
    trace, results = oracle.build_fib_trace(1, 256)
    tm = np.array([[oracle.to_mont(int(v)) for v in row] for row in trace], dtype=np.uint64)
    opts = oracle.make_opts(folding=8, rem_max_deg=31, grinding=2)
    assert ctx.prove_fib(tm, results, opts, mont=True) == oracle.prove_fib(trace, results, opts)


@pytest.mark.parametrize("h,g", [(wf.HASH_BLAKE3_256, 16), (wf.HASH_BLAKE3_256, 21), (wf.HASH_RP64_256, 10), (wf.HASH_RPJIVE64_256, 9), (wf.HASH_BLAKE3_192, 14), (wf.HASH_SHA3_256, 11)])
def test_grinding_smallest_nonce(ctx, oracle, h, g):
    # prover/src/channel.rs:173-175 (serial branch): smallest nonce
    coin = oracle.RandomCoin(h, [1, 2, 3, g])
    want = 1
    while coin.leading_zeros(want) < g:
        want += 1
    assert ctx.grind(h, coin.seed, g) == want


# ---- generic AIR front end (wf_prove_air): same description drives oracle prover, oracle verifier, device ----
import airs  # noqa: E402  (tests/airs.py)


@pytest.mark.parametrize("k,log_n,ext", [(1, 7, 1), (4, 9, 3)])
def test_generic_air_reproduces_fib_path(ctx, oracle, k, log_n, ext):
    # the FibSmall x k AIR written as a program must give exactly the specialised kernel's proof
    desc, trace = airs.fib_small_x(k, 1 << log_n)
    _, results = oracle.build_fib_trace(k, 1 << log_n)
    opts = oracle.make_opts(ext=ext, grinding=3, folding=4, rem_max_deg=7)
    assert ctx.prove_air(desc, trace, opts) == ctx.prove_fib(trace, results, opts)


@pytest.mark.parametrize("name,log_n", [("mulfib2", 8), ("periodic_mix", 9), ("periodic_mix", 12)])
@pytest.mark.parametrize("ext,h,batch", [(1, wf.HASH_BLAKE3_256, 0), (3, wf.HASH_BLAKE3_256, 1), (2, wf.HASH_RP64_256, 2)])
def test_generic_air_vs_oracle(ctx, oracle, name, log_n, ext, h, batch):
    # mulfib2 = examples/src/fibonacci/mulfib2/air.rs over f64 (degree 2); periodic_mix exercises periodic
    # columns, a periodic assertion, degree 3 (ce_blowup 4, three composition columns) and two exemptions
    desc, trace = getattr(airs, name)(1 << log_n)
    opts = oracle.make_opts(num_queries=24, blowup=8, grinding=2, ext=ext, folding=4, rem_max_deg=15, batch_c=batch, batch_d=batch, hash_id=h)
    got = ctx.prove_air(desc, trace, opts)
    assert got == oracle.prove_air(desc, trace, opts)
    assert oracle.verify_air(desc, got, h) == 0
    bad = desc.copy()
    bad[-2] ^= np.uint64(1)  # last public input
    assert oracle.verify_air(bad, got, h) != 0


# ---- auxiliary trace segment (wf_prove_air_aux): RAP-style two-segment AIR ----
@pytest.mark.parametrize("log_n", [6, 10])
@pytest.mark.parametrize("ext,h,batch", [(1, wf.HASH_BLAKE3_256, 0), (2, wf.HASH_BLAKE3_256, 1), (3, wf.HASH_RP64_256, 2)])
def test_aux_segment_vs_oracle(ctx, oracle, log_n, ext, h, batch):
    # the flow of examples/src/rescue_raps (prover/src/lib.rs:309-349): main commit -> draw random elements ->
    # host builds the aux columns over E -> aux commit -> constraints over both segments
    desc, trace, builder = airs.perm_rap(1 << log_n)
    opts = oracle.make_opts(num_queries=24, blowup=8, grinding=2, ext=ext, folding=4, rem_max_deg=7, batch_c=batch, batch_d=batch, hash_id=h)
    got = ctx.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert got == oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert oracle.verify_air(desc, got, h) == 0
    # single-segment entry point refuses a multi-segment description
    with pytest.raises(wf.WfError):
        ctx.prove_air(desc, trace, opts)


def test_aux_segment_montgomery_io(ctx, oracle):
    # mont=1: trace, random elements and aux columns all cross the ABI as the reference's Montgomery words
    desc, trace, builder = airs.perm_rap(64)
    opts = oracle.make_opts(num_queries=16, blowup=8, ext=2, folding=4, rem_max_deg=7)
    to_m = np.vectorize(lambda v: oracle.to_mont(int(v)), otypes=[np.uint64])
    from_m = np.vectorize(lambda v: oracle.from_mont(int(v)), otypes=[np.uint64])
    got = ctx.prove_air_aux(desc, to_m(trace), opts, lambda r: to_m(builder(from_m(r))), airs.PERM_RAP_AUX_WIDTH, 2, mont=True)
    assert got == oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)


def test_aux_segment_inconsistent_trace_is_rejected(ctx, oracle):
    desc, trace, builder = airs.perm_rap(128)
    bad = trace.copy()
    bad[2, 7] = (int(bad[2, 7]) + 1) % airs.P  # b is no longer a permutation of x0
    opts = oracle.make_opts(num_queries=16, blowup=8, ext=2, folding=4, rem_max_deg=7)
    proof = ctx.prove_air_aux(desc, bad, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert oracle.verify_air(desc, proof) != 0


# ---- sequence assertions (SmallPoly / LargePoly boundary constraints) ----
@pytest.mark.parametrize("log_n", [6, 11])
@pytest.mark.parametrize("ext,h,batch", [(1, wf.HASH_BLAKE3_256, 0), (3, wf.HASH_BLAKE3_256, 2)])
def test_sequence_assertions_vs_oracle(ctx, oracle, log_n, ext, h, batch):
    # n / 4 asserted values (16: the reference's Horner path; 512: its pre-evaluated table path,
    # prover/src/constraints/evaluator/boundary.rs:340,389) with first_step = 1, sharing a divisor with a
    # periodic assertion, plus a two-value sequence at first_step 0
    desc, trace = airs.sequence_mix(1 << log_n)
    opts = oracle.make_opts(num_queries=20, blowup=8, grinding=1, ext=ext, folding=4, rem_max_deg=7, batch_c=batch, batch_d=batch, hash_id=h)
    got = ctx.prove_air(desc, trace, opts)
    assert got == oracle.prove_air(desc, trace, opts)
    assert oracle.verify_air(desc, got, h) == 0
    # a description asserting a different sequence value does not verify this proof
    bad = desc.copy()
    i = int(np.nonzero(bad == np.uint64(trace[0, 5]))[0][0])  # trace[0, 5] is the second value of the sequence
    bad[i] ^= np.uint64(1)
    assert oracle.verify_air(bad, got, h) != 0


def test_invalid_sequence_is_refused(ctx, oracle):
    desc, trace = airs.sequence_mix(64)
    opts = oracle.make_opts(num_queries=8, blowup=8)
    with pytest.raises(wf.WfError):
        ctx.prove_air(desc, trace[:, :32].copy(), opts)  # n / stride no longer equals the number of values


# ---- the corners of ProofOptions::new (air/src/options.rs:132-190) and TraceInfo (trace_info.rs:60-110) ----
EDGE_CASES = [
    (1, 3, dict(num_queries=5, blowup=128, grinding=0, ext=1, folding=16, rem_max_deg=7)),     # N = (7 + 1) * 128: no FRI layer at all
    (1, 3, dict(num_queries=255, blowup=2, grinding=0, ext=3, folding=2, rem_max_deg=0)),      # minimum length / blowup / remainder, maximum queries
    (127, 4, dict(num_queries=16, blowup=4, grinding=3, ext=2, folding=4, rem_max_deg=1)),     # 254 columns (MAX_TRACE_WIDTH = 255)
    (2, 10, dict(num_queries=64, blowup=16, grinding=0, ext=1, folding=16, rem_max_deg=255)),  # maximum remainder degree
    (1, 6, dict(num_queries=1, blowup=8, grinding=0, ext=1, folding=8, rem_max_deg=3, batch_c=1, batch_d=2)),
    (3, 5, dict(num_queries=40, blowup=64, grinding=0, ext=2, folding=2, rem_max_deg=15, hash_id=1)),
]


@pytest.mark.parametrize("k,log_n,kw", EDGE_CASES)
def test_option_corners_match_oracle(ctx, oracle, k, log_n, kw):
    trace, results = oracle.build_fib_trace(k, 1 << log_n)
    opts = oracle.make_opts(**kw)
    got = ctx.prove_fib(trace, results, opts)
    assert got == oracle.prove_fib(trace, results, opts)
    assert oracle.verify_fib(got, k, results, kw.get("hash_id", 0)) == 0


@pytest.mark.parametrize("kw", [dict(blowup=3), dict(blowup=256), dict(num_queries=0), dict(num_queries=256), dict(folding=3),
                                dict(ext=4), dict(hash_id=7), dict(grinding=33)])
def test_invalid_options_are_refused(ctx, oracle, kw):
    # the reference panics in ProofOptions::new; the C ABI returns an error instead of proving
    trace, results = oracle.build_fib_trace(1, 64)
    base = dict(num_queries=8, blowup=8, grinding=0, ext=1, folding=4, rem_max_deg=7)
    base.update(kw)
    with pytest.raises(wf.WfError):
        ctx.prove_fib(trace, results, oracle.make_opts(**base))


# ---- committed golden fixtures: the device pipeline against tests/golden/proof_digests.json ----
@pytest.mark.parametrize("idx", range(11))
def test_device_reproduces_golden_proof_digests(ctx, oracle, idx):
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proof_digests.json")) as f:
        rec = json.load(f)[idx]
    opts = oracle.make_opts(**rec["opts"])
    if rec["kind"] == "fib":
        trace, res = oracle.build_fib_trace(rec["k"], 1 << rec["log_n"])  # trace builder only: the proof is the product's
        proof = ctx.prove_fib(trace, res, opts)
    elif rec["air"] == "perm_rap":
        desc, trace, builder = airs.perm_rap(rec["n"])
        proof = ctx.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    else:
        desc, trace = getattr(airs, rec["air"])(rec["n"])
        proof = ctx.prove_air(desc, trace, opts)
    assert len(proof) == rec["bytes"] and hashlib.sha256(proof).hexdigest() == rec["sha256"]


def test_overlapping_assertions_are_refused(ctx, oracle):
    # the reference panics in prepare_assertions when two assertions cover the same cell
    # (air/src/air/boundary/mod.rs:205-210, Assertion::overlaps_with assertions/mod.rs:175-208)
    n = 64
    desc, trace = airs.mulfib2(n)
    A = airs.AirBuilder(2)
    A.pub = [int(trace[0, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.mul(A.cur(0), A.cur(1))), 2)
    A.constraint(A.sub(A.nxt(1), A.mul(A.cur(1), A.nxt(0))), 2)
    A.assert_single(0, 0, 1)
    A.assert_single(1, 0, 2)
    A.assert_sequence(1, 0, n // 2, [2, int(trace[1, n // 2])])   # step 0 of column 1 is asserted twice
    opts = oracle.make_opts(num_queries=8, blowup=8)
    with pytest.raises(wf.WfError, match="overlaps"):
        ctx.prove_air(A.build(), trace, opts)
    B = airs.AirBuilder(2)
    B.pub = A.pub
    B.constraint(B.sub(B.nxt(0), B.mul(B.cur(0), B.cur(1))), 2)
    B.constraint(B.sub(B.nxt(1), B.mul(B.cur(1), B.nxt(0))), 2)
    B.assert_single(0, 0, 1)
    B.assert_periodic(1, 0, 4, 2)
    B.assert_single(1, 8, 5)                                         # step 8 = 0 + 2 * 4 is covered by the periodic one
    with pytest.raises(wf.WfError, match="overlaps"):
        ctx.prove_air(B.build(), trace, opts)


# ProofOptions::with_partitions (air/src/options.rs:193-200) through the one-call entry points: main / aux / constraint rows
# hashed as merge_many over column partitions (row_matrix.rs:204-223). (G, hash_rate) = (2,8), (4,8), (8,8) are the
# multi-GPU-motivated settings of SURVEY.md 8e option 2; (4, 64) makes partition_size exceed the row width (one partition,
# still merge_many); (16, 1) is the maximum partition count.
@pytest.mark.parametrize("k,log_n,ext,h,parts,rate", [
    (4, 10, 1, wf.HASH_BLAKE3_256, 2, 8), (4, 10, 3, wf.HASH_BLAKE3_256, 4, 8), (32, 9, 3, wf.HASH_BLAKE3_256, 8, 8),
    (8, 9, 2, wf.HASH_RP64_256, 2, 8), (4, 9, 1, wf.HASH_RP64_256, 8, 8), (4, 8, 3, wf.HASH_RPJIVE64_256, 4, 4), (8, 9, 3, wf.HASH_BLAKE3_192, 4, 8), (16, 9, 1, wf.HASH_SHA3_256, 2, 8), (1, 9, 3, wf.HASH_BLAKE3_256, 4, 64),
    (16, 9, 1, wf.HASH_BLAKE3_256, 16, 1)])
def test_partitioned_commitments_match_oracle(ctx, oracle, k, log_n, ext, h, parts, rate):
    trace, results = oracle.build_fib_trace(k, 1 << log_n)
    opts = oracle.make_opts(num_queries=24, grinding=3, ext=ext, folding=4, rem_max_deg=7, hash_id=h, num_partitions=parts, hash_rate=rate)
    want = oracle.prove_fib(trace, results, opts)
    got = ctx.prove_fib(trace, results, opts)
    assert got == want
    assert oracle.verify_fib(got, k, results, h) == 0
    assert got != ctx.prove_fib(trace, results, oracle.make_opts(num_queries=24, grinding=3, ext=ext, folding=4, rem_max_deg=7, hash_id=h))


@pytest.mark.parametrize("ext,h,parts,rate", [(2, wf.HASH_BLAKE3_256, 2, 2), (3, wf.HASH_RP64_256, 3, 8), (2, wf.HASH_RPJIVE64_256, 2, 4)])
def test_partitioned_aux_segment_vs_oracle(ctx, oracle, ext, h, parts, rate):
    # the auxiliary commitment uses partition_size::<E>(aux_width) (trace_lde/default/mod.rs:147), the constraint commitment
    # partition_size::<E>(num composition columns) (constraints/commitment/default.rs:147)
    desc, trace, builder = airs.perm_rap(1 << 8)
    opts = oracle.make_opts(num_queries=20, grinding=2, ext=ext, folding=4, rem_max_deg=7, hash_id=h, num_partitions=parts, hash_rate=rate)
    got = ctx.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert got == oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    assert oracle.verify_air(desc, got, h) == 0


def test_too_many_partitions_refused(ctx, oracle):
    trace, results = oracle.build_fib_trace(1, 256)
    opts = oracle.make_opts(num_partitions=17, hash_rate=8)   # PartitionOptions::new asserts <= 16 (air/src/options.rs:413-414)
    with pytest.raises(wf.WfError):
        ctx.prove_fib(trace, results, opts)


@pytest.mark.parametrize("ext,h,mont", [(1, wf.HASH_BLAKE3_256, False), (3, wf.HASH_BLAKE3_256, False), (2, wf.HASH_RP64_256, True)])
def test_aux_assertions_depending_on_random_elements(ctx, oracle, ext, h, mont):
    # wf_prove_air_aux_dyn: Air::get_aux_assertions(aux_rand_elements) (air/src/air/mod.rs:279) as a host callback on the drawn
    # random elements; bytes equal to the oracle's prover given the same callback, accepted by the oracle's verifier with it
    n = 1 << 8
    desc, trace, builder = airs.perm_rap(n, dyn_last_q=True)
    opts = oracle.make_opts(num_queries=20, grinding=2, ext=ext, folding=4, rem_max_deg=7, hash_id=h)
    nv = builder.num_values
    want = oracle.prove_air_aux_dyn(desc, trace, opts, builder, builder.values_fn, airs.PERM_RAP_AUX_WIDTH, 2, nv)
    if mont:   # every word crossing the ABI (trace, random elements, aux columns, assertion values) in Montgomery form
        to_m = np.vectorize(lambda v: oracle.to_mont(int(v)), otypes=[np.uint64])
        from_m = np.vectorize(lambda v: int(v) * pow(2**64, -1, wf.P) % wf.P, otypes=[np.uint64])
        b = lambda rand: to_m(builder(from_m(rand)))
        vf = lambda rand, values: to_m(builder.values_fn(from_m(rand), from_m(values)))
        got = ctx.prove_air_aux_dyn(desc, to_m(trace), opts, b, vf, airs.PERM_RAP_AUX_WIDTH, 2, nv, mont=True)
    else:
        got = ctx.prove_air_aux_dyn(desc, trace, opts, builder, builder.values_fn, airs.PERM_RAP_AUX_WIDTH, 2, nv)
    assert got == want
    assert oracle.verify_air_dyn(desc, got, h, builder.values_fn, 2, nv, ext) == 0
    assert oracle.verify_air(desc, got, h) != 0      # the description's placeholder is not the asserted value


@pytest.mark.parametrize("name,log_n,ext", [("rescue_like", 10, 1), ("rescue_like", 9, 3), ("periodic_mix", 10, 2), ("sequence_mix", 9, 3)])
def test_compiled_constraint_kernel_equals_interpreter_and_oracle(ctx, oracle, name, log_n, ext):
    # the per-AIR kernel compiled with NVRTC (jit.cu; registers and constants as literals) and the built-in interpreter must
    # produce the same proof bytes, equal to the oracle's; rescue_like is the heavy case (degree-7 S-boxes, MDS layer)
    desc, trace = getattr(airs, name)(1 << log_n)
    opts = oracle.make_opts(num_queries=20, blowup=8, grinding=2, ext=ext, folding=4, rem_max_deg=7)
    want = oracle.prove_air(desc, trace, opts)
    before = ctx.jit_stats()
    ctx.set_jit(True)
    got_jit = ctx.prove_air(desc, trace, opts)
    after = ctx.jit_stats()
    ctx.set_jit(False)
    got_int = ctx.prove_air(desc, trace, opts)
    ctx.set_jit(True)
    assert got_jit == want and got_int == want
    assert oracle.verify_air(desc, got_jit, 0) == 0
    # the compiled kernel really ran (no silent fallback to the interpreter)
    assert after["fallbacks"] == before["fallbacks"]
    assert after["compiled"] + after["cache_hits"] > before["compiled"] + before["cache_hits"]


def test_compiled_constraint_kernel_aux_segment(ctx, oracle):
    desc, trace, builder = airs.perm_rap(1 << 9)
    opts = oracle.make_opts(num_queries=20, grinding=2, ext=3, folding=4, rem_max_deg=7)
    want = oracle.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2)
    s0 = ctx.jit_stats()
    assert ctx.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2) == want
    s1 = ctx.jit_stats()
    assert s1["fallbacks"] == s0["fallbacks"] and s1["compiled"] + s1["cache_hits"] > s0["compiled"] + s0["cache_hits"]
    ctx.set_jit(False)
    assert ctx.prove_air_aux(desc, trace, opts, builder, airs.PERM_RAP_AUX_WIDTH, 2) == want
    ctx.set_jit(True)


def test_no_device_buffers_left_behind(oracle):
    # every proof entry point — succeeding or refusing its input — hands all of its device buffers back to the context's
    # pool (wf_ctx_mem_stats: nothing live once the caller holds no handle)
    c = wf.Context(0)
    try:
        trace, results = oracle.build_fib_trace(4, 1 << 10)
        opts = oracle.make_opts(ext=3, grinding=2, folding=4, rem_max_deg=7)
        c.prove_fib(trace, results, opts)
        desc, tr = airs.sequence_mix(1 << 8)
        c.prove_air(desc, tr, oracle.make_opts(num_queries=8, blowup=8, ext=2))
        with pytest.raises(wf.WfError):
            c.prove_air(desc, tr[:, :128].copy(), oracle.make_opts(num_queries=8, blowup=8))
        d2, t2, builder = airs.perm_rap(128)
        c.prove_air_aux(d2, t2, oracle.make_opts(num_queries=16, blowup=8, ext=2, folding=4, rem_max_deg=7), builder, airs.PERM_RAP_AUX_WIDTH, 2)
        with pytest.raises(wf.WfError):
            c.prove_fib(trace, results, oracle.make_opts(num_partitions=17, hash_rate=8))
        m = c.mat_from_host_columns(oracle.rand_elems((1, 1 << 12), 5))
        cw = m.lde(3)
        assert c.mem_stats()[0] == 2           # the two matrices the test holds
        f, _ = c.fri_build_layers_default(wf.HASH_BLAKE3_256, cw, 1, 4, 7, 8)
        assert c.mem_stats()[0] > 2            # + the FRI layers and their trees
        f.free(); m.free(); cw.free()
        live, live_bytes, pooled = c.mem_stats()
        assert live == 0 and live_bytes == 0 and pooled > 0
    finally:
        c.close()
