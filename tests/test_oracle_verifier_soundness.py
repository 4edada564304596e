"""The oracle's verifier is what the GPU proofs are accepted by (tests/test_gpu_*.py), so it must not be lenient: every
single-bit change of a valid proof — context, options, commitments, openings, Merkle paths, OOD frame, FRI layers,
remainder, FriProof::num_partitions, nonce — has to be refused, as do truncated and extended proofs. The only byte the
reference itself does not bind is ProofOptions' hash_rate while num_partitions = 1 (air/src/options.rs:428-431: unused,
and not part of Context::to_elements); zero there is refused like PartitionOptions::new does (:416)."""
import numpy as np
import pytest

HASH_RATE_BYTE = 24   # context: 4 + 2 (no meta) + 1 + 8 modulus = 15 bytes, then the ten option bytes; hash_rate is the last


@pytest.mark.parametrize("h,ext", [(0, 1), (1, 2), (3, 3)])
def test_every_byte_of_a_proof_is_bound(oracle, h, ext):
    k, n = 1, 128
    trace, results = oracle.build_fib_trace(k, n)
    opts = oracle.make_opts(num_queries=12, blowup=8, grinding=2, ext=ext, folding=4, rem_max_deg=7, hash_id=h)
    proof = oracle.prove_fib(trace, results, opts)
    assert oracle.verify_fib(proof, k, results, h) == 0
    rng = np.random.default_rng(h * 10 + ext)
    accepted = []
    L = len(proof)
    # every byte of the context / options / commitments head and of the remainder / nonce tail, every fourth byte in between
    # (openings, Merkle paths, FRI layers: thousands of bytes of the same kind)
    for i in sorted(set(range(0, 160)) | set(range(L - 160, L)) | set(range(160, L - 160, 4))):
        b = bytearray(proof)
        b[i] ^= 1 << int(rng.integers(0, 8))
        if oracle.verify_fib(bytes(b), k, results, h) == 0:
            accepted.append(i)
    assert set(accepted) <= {HASH_RATE_BYTE}, accepted
    zero_rate = bytearray(proof)
    zero_rate[HASH_RATE_BYTE] = 0
    assert oracle.verify_fib(bytes(zero_rate), k, results, h) != 0
    for cut in (0, 1, 14, 15, 25, len(proof) // 2, len(proof) - 8, len(proof) - 1):
        assert oracle.verify_fib(proof[:cut], k, results, h) != 0
    assert oracle.verify_fib(proof + b"\x00", k, results, h) != 0
    # the trace length travels as log2 in one byte: 7 + 64 must not be read as 2^7 again
    wrap = bytearray(proof)
    wrap[3] += 64
    assert oracle.verify_fib(bytes(wrap), k, results, h) != 0
