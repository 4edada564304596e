"""CPU-side tests of the product: the C-ABI library loads, exports every symbol the header declares,
and its host-side transcript arithmetic agrees with the oracle. No GPU, no compute kernels."""
import os
import re

import numpy as np
import pytest

import winterfell_b200 as wf

P = wf.P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "winterfell_b200.h")).read()
    declared = set(re.findall(r"\b(wf_[a-z0-9_]+)\s*\(", hdr)) - {"wf_fri_commit_fn", "wf_fri_draw_fn"}
    L = wf.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(wf.declared_symbols()), declared ^ set(wf.declared_symbols())
    assert b"sm_100a" in L.wf_version()


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(wf.WfError):
        wf.Context(0)


def test_host_field_helpers(oracle):
    L = wf.lib()
    rng = np.random.default_rng(3)
    xs = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0xFFFFFFFF00000000, 1 << 63] + [int(v) % P for v in rng.integers(0, 2**64, 64, dtype=np.uint64)]
    for x in xs:
        for k in (0, 3, 6, 12, 24, 36, 48, 60, 63, 64, 65, 72, 84, 95, 96):
            assert L.wf_host_mul_2exp(x, k) == x * pow(2, k, P) % P, (x, k)
        assert L.wf_host_mont_to_canonical(oracle.to_mont(x)) == x
        assert L.wf_host_canonical_to_mont(x) == oracle.to_mont(x)
    for a, b in zip(xs, reversed(xs)):
        assert L.wf_host_mul(a, b) == a * b % P


@pytest.mark.parametrize("h", [wf.HASH_BLAKE3_256, wf.HASH_RP64_256])
def test_host_hashers_match_oracle(oracle, h):
    for n in (0, 1, 4, 7, 8, 9, 16, 100, 128, 129, 300, 1000):
        e = oracle.rand_elems(n, 100 + n)
        assert wf.host_hash_elements(h, e) == oracle.hash_elements(h, e), n
    a = oracle.hash_elements(h, [1, 2, 3])
    b = oracle.hash_elements(h, [4, 5])
    assert wf.host_merge(h, a, b) == oracle.merge(h, a, b)
    for v in (0, 5, P - 1, P, P + 2, 2**64 - 1):
        assert wf.host_merge_with_int(h, a, v) == oracle.merge_with_int(h, a, v), v


def test_host_usize_encoding_matches_reference_test():
    # utils/core/src/tests.rs::write_serializable_usize, through the product's serializer
    import ctypes as C
    L = wf.lib()
    buf = (C.c_uint8 * 9)()
    total = 0
    for v, want_total in [(0, 1), (1, 2), (255, 4), (234567, 7), (2**64 - 1, 16)]:
        ln = L.wf_host_write_usize(v, buf)
        total += ln
        assert total == want_total
        b = bytes(buf[:ln])
        length = 9 if b[0] == 0 else ((b[0] & -b[0]).bit_length())
        assert length == ln
        assert (int.from_bytes(b[1:9], "little") if length == 9 else int.from_bytes(b, "little") >> length) == v


@pytest.mark.parametrize("h", [0, 1])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_host_coin_matches_oracle(oracle, h, d):
    # the product's DefaultRandomCoin (host_transcript.hpp) against the oracle's (crypto/src/random/default.rs)
    import ctypes as C
    L = wf.lib()
    seed = oracle.rand_elems((7,), 3 + d)
    root = bytes(range(32))
    out = np.zeros((20, d), dtype=np.uint64)
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    rb = (C.c_uint8 * 32)(*root)
    assert L.wf_host_coin_draw(h, seed.ctypes.data_as(u64p), seed.size, rb, d, 20, out.ctypes.data_as(u64p)) == 0
    coin = oracle.RandomCoin(h, seed)
    coin.reseed(root)
    for i in range(20):
        assert (coin.draw(d) == out[i]).all()
