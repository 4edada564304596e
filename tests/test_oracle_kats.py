"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for this
path. Vectors are cited by reference file:line. Runs on CPU (`-m "not gpu"`)."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def naive_eval(p, x):
    acc = 0
    for c in reversed([int(v) for v in p]):
        acc = (acc * x + c) % P
    return acc


# ---- field: math/src/field/f64/tests.rs ----
def test_field_basics(oracle):
    o = oracle
    assert o.mul(P - 1, P - 1) == 1                      # tests.rs:54-74 (m-1)^2 = 1
    assert o.mul(o.inv(2), 2) == 1
    assert o.add(P - 1, 1) == 0 and o.sub(0, 1) == P - 1
    assert o.inv(0) == 0
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = (int(v) % P for v in rng.integers(0, 2**64, size=2, dtype=np.uint64))
        assert o.mul(a, b) == a * b % P
        assert o.add(a, b) == (a + b) % P
        assert o.sub(a, b) == (a - b) % P
    assert o.exp(7, P - 1) == 1
    # root of unity order (tests.rs get_root_of_unity): g^(2^32) = 1, g^(2^31) != 1
    g = o.root_of_unity(32)
    assert g == 7277203076849721926
    assert o.exp(g, 1 << 32) == 1 and o.exp(g, 1 << 31) == P - 1
    for k in range(1, 7):  # small roots are powers of two (used by the device kernels)
        assert o.root_of_unity(k) == pow(2, 192 >> k, P) if k <= 6 else True


def test_montgomery_view(oracle):
    # f64/tests.rs:173-189: elements_as_bytes exposes Montgomery words x * 2^64 mod p
    for x in (0, 1, 2, 7, P - 1, 123456789123456789):
        m = oracle.to_mont(x)
        assert m == (x << 64) % P
        assert oracle.from_mont(m) == x


def test_quad_mul_kats(oracle):
    # f64/tests.rs:220-246
    m = P
    cases = [((3, 1), (4, 2), (8, 12)),
             ((3, m - 1), (m - 3, 5), (1, 13)),
             ((3, m - 1), (10, m - 2), (26, 18446744069414584307))]
    for a, b, e in cases:
        assert list(oracle.ext_mul(a, b)) == list(e)


def test_cube_mul_kats(oracle):
    # f64/tests.rs:288-345
    cases = [((3, 5, 2), (320, 68, 3), (1111, 1961, 995)),
             ((18446744069414584267, 18446744069414584309, 9223372034707292160),
              (18446744069414584101, 420, 18446744069414584121),
              (14070, 18446744069414566571, 5970)),
             ((18446744069414584266, 18446744069412558094, 5268562),
              (18446744069414583589, 1226, 5346),
              (18446744065041672051, 25275910656, 21824696736))]
    for a, b, e in cases:
        assert list(oracle.ext_mul(a, b)) == list(e)


@pytest.mark.parametrize("d", [2, 3])
def test_ext_inv(oracle, d):
    a = oracle.rand_elems(d, 5)
    one = np.zeros(d, dtype=np.uint64); one[0] = 1
    assert list(oracle.ext_mul(a, oracle.ext_inv(a))) == list(one)
    assert list(oracle.ext_inv(np.zeros(d, dtype=np.uint64))) == [0] * d


# ---- fft: math/src/fft/tests.rs ----
def test_twiddles(oracle):
    # tests.rs:63-73: twiddles == bit-reversed power series of the n-th root
    n = 32
    g = oracle.root_of_unity(5)
    tw = oracle.get_twiddles(n)
    for i in range(n // 2):
        j = int(format(i, "04b")[::-1], 2)
        assert int(tw[j]) == pow(g, i, P)
    itw = oracle.get_inv_twiddles(n)
    for i in range(n // 2):
        assert oracle.mul(int(tw[i]), int(itw[i])) == 1


@pytest.mark.parametrize("n", [4, 8, 16, 1024])
def test_fft_vs_naive(oracle, n):
    # tests.rs:20-61: evaluate_poly == polynom::eval_many over the n-th roots; interpolate inverts it
    p = oracle.rand_elems(n, n)
    ev = oracle.evaluate_poly(p)
    g = oracle.root_of_unity(n.bit_length() - 1)
    pts = range(n) if n <= 16 else [0, 1, 5, n // 2, n - 1]
    for i in pts:
        assert int(ev[i]) == naive_eval(p, pow(g, i, P))
    assert (oracle.interpolate_poly(ev) == p).all()


@pytest.mark.parametrize("d", [1, 2, 3])
def test_fft_with_offset(oracle, d):
    # fft/mod.rs doc-tests :144-167, :328-350
    n, b = 16, 4
    p = oracle.rand_elems(n * d, 3)
    ev = oracle.evaluate_poly_with_offset(p, 7, b, d)
    g = oracle.root_of_unity(6)
    for comp in range(d):
        pc = p[comp::d]
        for i in (0, 1, 17, 63):
            assert int(ev[i * d + comp]) == naive_eval(pc, 7 * pow(g, i, P) % P)
    # interpolate_with_offset inverts evaluation over the shifted size-n domain
    ev1 = oracle.evaluate_poly_with_offset(p, 7, 1, d)
    assert (oracle.interpolate_poly_with_offset(ev1, 7, d) == p).all()


def test_lde_matrix_vs_naive(oracle):
    # prover/src/matrix/tests.rs:15-42: f64, 64 columns, n=256, blowup 8, row-major == naive eval
    n, c, b = 256, 64, 8
    polys = oracle.rand_elems((c, n), 42)
    lde = oracle.lde_rows(polys, b)
    assert lde.shape == (n * b, c)
    g = oracle.root_of_unity(11)
    for row in (0, 1, 7, 8, 1000, n * b - 1):
        x = 7 * pow(g, row, P) % P
        for col in (0, 1, 7, 8, 33, 63):
            assert int(lde[row, col]) == naive_eval(polys[col], x)
    # every column equals evaluate_poly_with_offset of that column
    for col in (0, 13, 63):
        assert (lde[:, col] == oracle.evaluate_poly_with_offset(polys[col], 7, b)).all()


def test_trace_lde_reinterpolates(oracle):
    # prover/src/trace/trace_lde/default/tests.rs:22-72: interpolate trace -> polys evaluate to the
    # Fibonacci trace; LDE of the polys re-interpolates to the same polys
    n, b = 64, 8
    tr = np.zeros((2, n), dtype=np.uint64)
    a0, a1 = 1, 1
    for i in range(n):
        tr[0, i], tr[1, i] = a0, a1
        a0 = (a0 + a1) % P
        a1 = (a1 + a0) % P
    polys = oracle.interpolate_columns(tr)
    g = oracle.root_of_unity(6)
    for i in (0, 1, 2, 63):
        assert naive_eval(polys[0], pow(g, i, P)) == int(tr[0, i])
        assert naive_eval(polys[1], pow(g, i, P)) == int(tr[1, i])
    lde = oracle.lde_rows(polys, b)
    for col in range(2):
        back = oracle.interpolate_poly_with_offset(np.ascontiguousarray(lde[:, col]), 7)
        assert (back[:n] == polys[col]).all() and not back[n:].any()
    # commitment == MerkleTree(hash_elements(rows)) (tests.rs:74-106)
    dg = oracle.hash_rows(oracle.BLAKE3, lde)
    assert dg[5].tobytes() == oracle.hash_elements(oracle.BLAKE3, lde[5])


# ---- hashers ----
def test_blake3_vs_spec(oracle):
    import blake3 as b3
    rng = np.random.default_rng(7)
    for ln in (0, 1, 31, 32, 40, 63, 64, 65, 127, 128, 512, 1023, 1024, 1025, 2048, 2049, 3072, 4097, 8192 + 17):
        data = rng.integers(0, 256, size=ln, dtype=np.uint8).tobytes()
        assert oracle.blake3(data) == b3.blake3(data).digest(), ln


def test_blake3_hasher_semantics(oracle):
    import blake3 as b3
    e = oracle.rand_elems(9, 11)
    # blake/mod.rs:52-65: canonical LE bytes, no length prefix
    assert oracle.hash_elements(oracle.BLAKE3, e) == b3.blake3(e.astype("<u8").tobytes()).digest()
    a, b = bytes(range(32)), bytes(range(32, 64))
    assert oracle.merge(oracle.BLAKE3, a, b) == b3.blake3(a + b).digest()          # :33
    assert oracle.merge_with_int(oracle.BLAKE3, a, 2**63 + 5) == b3.blake3(a + (2**63 + 5).to_bytes(8, "little")).digest()  # :41


def test_rp64_permutation_kat(oracle):
    # crypto/src/hash/rescue/rp64_256/tests.rs:69-105 (Sage reference vector)
    expected = [11084501481526603421, 6291559951628160880, 13626645864671311919, 18397438323058963117,
                7443014167353970324, 17930833023906771425, 4275355080008025761, 7676681476902901785,
                3460534574143792217, 11912731278641497187, 8104899243369883110, 674509706691634438]
    assert [int(v) for v in oracle.rp64_permute(np.arange(12, dtype=np.uint64))] == expected


def test_rp64_consistency(oracle):
    # rp64_256/tests.rs:107-159
    R = oracle.RP64
    e = oracle.rand_elems(8, 21)
    m = oracle.merge(R, e[:4].tobytes(), e[4:].tobytes())
    assert m == oracle.hash_elements(R, e)
    assert m == oracle.merge_many(R, e.tobytes())
    seed = oracle.rand_elems(4, 22)
    v = int(oracle.rand_elems(1, 23)[0])
    assert oracle.merge_with_int(R, seed.tobytes(), v) == oracle.hash_elements(R, np.array([int(x) for x in seed] + [v], dtype=np.uint64))
    v = P + 2
    assert oracle.merge_with_int(R, seed.tobytes(), v) == oracle.hash_elements(R, np.array([int(x) for x in seed] + [v % P, 1], dtype=np.uint64))


def test_rpjive_permutation_kat(oracle):
    # crypto/src/hash/rescue/rp64_256_jive/tests.rs:69-101 (Sage reference vector)
    expected = [16940713730596720799, 16218555904323712189, 11042680722444601138, 5370396747047489939,
                6349480890410006944, 1551053614279730715, 3995941143622927528, 9350074312471431779]
    assert [int(v) for v in oracle.rpjive_permute(np.arange(8, dtype=np.uint64))] == expected


def _py_rpjive_hash_elements(oracle, elems):
    # rp64_256_jive/mod.rs:240-282 restated on Python integers over the (KAT-pinned) permutation
    s = [0] * 8
    if len(elems) % 4:
        s[0] = 1
    i = 0
    for e in elems:
        s[4 + i] = (s[4 + i] + int(e)) % P
        i += 1
        if i == 4:
            s = [int(v) for v in oracle.rpjive_permute(np.array(s, dtype=np.uint64))]
            i = 0
    if i > 0:
        s[4 + i] = 1
        for q in range(i + 1, 4):
            s[4 + q] = 0
        s = [int(v) for v in oracle.rpjive_permute(np.array(s, dtype=np.uint64))]
    return np.array(s[4:8], dtype=np.uint64).tobytes()


def _py_rpjive_compress(oracle, state):
    out = [int(v) for v in oracle.rpjive_permute(np.array(state, dtype=np.uint64))]
    return np.array([(state[i] + state[4 + i] + out[i] + out[4 + i]) % P for i in range(4)], dtype=np.uint64).tobytes()


def test_rpjive_hasher_structure(oracle):
    # the sponge (padding by overwriting, capacity flag), the Jive compression of merge / merge_with_int
    # (rp64_256_jive/mod.rs:186-229, 337-350) and merge_many = hash_elements (:198-200), against a Python restatement; and the
    # reference's own consistency tests (tests.rs:103-140): merge and merge_with_int must DIFFER from hash_elements
    J = oracle.RPJIVE
    for n in (1, 3, 4, 5, 8, 9, 12, 17):
        e = oracle.rand_elems(n, 40 + n)
        assert oracle.hash_elements(J, e) == _py_rpjive_hash_elements(oracle, e)
    e = oracle.rand_elems(8, 21)
    m = oracle.merge(J, e[:4].tobytes(), e[4:].tobytes())
    assert m == _py_rpjive_compress(oracle, [int(v) for v in e])
    assert m != oracle.hash_elements(J, e)
    assert oracle.merge_many(J, e.tobytes()) == oracle.hash_elements(J, e)
    seed = oracle.rand_elems(4, 22)
    for v in (int(oracle.rand_elems(1, 23)[0]), P + 2, 0, P - 1, 2**64 - 1):
        st = [int(x) for x in seed] + [v % P, v // P if v >= P else 0, 0, 6 if v >= P else 5]
        got = oracle.merge_with_int(J, seed.tobytes(), v)
        assert got == _py_rpjive_compress(oracle, st)
        assert got != oracle.hash_elements(J, np.array([int(x) for x in seed] + [v % P], dtype=np.uint64))


def test_rpjive_round_trip_and_tamper(oracle):
    # a complete proof over RpJive64_256 (hash id 2): leaves by the sponge, tree nodes / coin by Jive compression
    k, n = 2, 128
    trace, res = oracle.build_fib_trace(k, n)
    for ext in (1, 2):
        opts = oracle.make_opts(ext=ext, hash_id=oracle.RPJIVE, grinding=3, folding=4, rem_max_deg=7)
        proof = oracle.prove_fib(trace, res, opts)
        assert oracle.verify_fib(proof, k, res, oracle.RPJIVE) == 0
        assert oracle.verify_fib(proof, k, res, oracle.RP64) != 0
        t = bytearray(proof)
        t[len(t) // 2] ^= 4
        assert oracle.verify_fib(bytes(t), k, res, oracle.RPJIVE) != 0


def test_blake3_192_against_the_spec(oracle):
    # Blake3_192 (crypto/src/hash/blake/mod.rs:73-123): the first 24 bytes of BLAKE3 over the element bytes / the 48 digest
    # bytes / seed bytes + u64; pinned to the public BLAKE3 spec through the Python package, digests in zero-padded slots
    import blake3 as b3
    H = oracle.BLAKE3_192
    e = oracle.rand_elems(13, 5)
    h = oracle.hash_elements(H, e)
    assert h[:24] == b3.blake3(e.tobytes()).digest()[:24] and h[24:] == bytes(8)
    a, b = oracle.hash_elements(H, e[:4]), oracle.hash_elements(H, e[4:9])
    assert oracle.merge(H, a, b) == b3.blake3(a[:24] + b[:24]).digest()[:24] + bytes(8)
    assert oracle.merge_many(H, a + b + a) == b3.blake3(a[:24] + b[:24] + a[:24]).digest()[:24] + bytes(8)
    for v in (0, 12345, 2**64 - 1):
        assert oracle.merge_with_int(H, a, v) == b3.blake3(a[:24] + v.to_bytes(8, "little")).digest()[:24] + bytes(8)
    trace, res = oracle.build_fib_trace(2, 128)
    for ext in (1, 3):
        opts = oracle.make_opts(ext=ext, hash_id=H, grinding=3, folding=4, rem_max_deg=7)
        proof = oracle.prove_fib(trace, res, opts)
        assert oracle.verify_fib(proof, 2, res, H) == 0 and oracle.verify_fib(proof, 2, res, oracle.BLAKE3) != 0
        assert len(proof) < len(oracle.prove_fib(trace, res, oracle.make_opts(ext=ext, hash_id=0, grinding=3, folding=4, rem_max_deg=7)))
        t = bytearray(proof)
        t[len(t) // 2] ^= 1
        assert oracle.verify_fib(bytes(t), 2, res, H) != 0


def test_sha3_256_against_the_spec(oracle):
    # Sha3_256 (crypto/src/hash/sha/mod.rs:19-60) pinned to FIPS 202 through hashlib; includes messages that end exactly on
    # the 136-byte rate (17 elements) and a complete proof
    import hashlib
    H = oracle.SHA3
    for n in (1, 4, 16, 17, 18, 34, 40):
        e = oracle.rand_elems(n, 70 + n)
        assert oracle.hash_elements(H, e) == hashlib.sha3_256(e.tobytes()).digest()
    a, b = oracle.hash_elements(H, oracle.rand_elems(3, 1)), oracle.hash_elements(H, oracle.rand_elems(5, 2))
    assert oracle.merge(H, a, b) == hashlib.sha3_256(a + b).digest()
    assert oracle.merge_many(H, a + b + a) == hashlib.sha3_256(a + b + a).digest()
    assert oracle.merge_with_int(H, a, 2**64 - 3) == hashlib.sha3_256(a + (2**64 - 3).to_bytes(8, "little")).digest()
    trace, res = oracle.build_fib_trace(2, 128)
    opts = oracle.make_opts(ext=3, hash_id=H, grinding=3, folding=4, rem_max_deg=7)
    proof = oracle.prove_fib(trace, res, opts)
    assert oracle.verify_fib(proof, 2, res, H) == 0 and oracle.verify_fib(proof, 2, res, oracle.BLAKE3) != 0


# ---- Merkle: crypto/src/merkle/tests.rs:14-84 ----
LEAVES4 = [
    [166, 168, 47, 140, 153, 86, 156, 86, 226, 229, 149, 76, 70, 132, 209, 109, 166, 193, 113, 197, 42, 116, 170, 144, 74, 104, 29, 110, 220, 49, 224, 123],
    [243, 57, 40, 140, 185, 79, 188, 229, 232, 117, 143, 118, 235, 229, 73, 251, 163, 246, 151, 170, 14, 243, 255, 127, 175, 230, 94, 227, 214, 5, 89, 105],
    [11, 33, 220, 93, 26, 67, 166, 154, 93, 7, 115, 130, 70, 13, 166, 45, 120, 233, 175, 86, 144, 110, 253, 250, 67, 108, 214, 115, 24, 132, 45, 234],
    [47, 173, 224, 232, 30, 46, 197, 186, 215, 15, 134, 211, 73, 14, 34, 216, 6, 11, 217, 150, 90, 242, 8, 31, 73, 85, 150, 254, 229, 244, 23, 231],
]
LEAVES8 = [
    [115, 29, 176, 48, 97, 18, 34, 142, 51, 18, 164, 235, 236, 96, 113, 132, 189, 26, 70, 93, 101, 143, 142, 52, 252, 33, 80, 157, 194, 52, 209, 129],
    [52, 46, 37, 214, 24, 248, 121, 199, 229, 25, 171, 67, 65, 37, 98, 142, 182, 72, 202, 42, 223, 160, 136, 60, 38, 255, 222, 82, 26, 27, 130, 203],
    [130, 43, 231, 0, 59, 228, 152, 140, 18, 33, 87, 27, 49, 190, 44, 82, 188, 155, 163, 108, 166, 198, 106, 143, 83, 167, 201, 152, 106, 176, 242, 119],
    [207, 158, 56, 143, 28, 146, 238, 47, 169, 32, 166, 97, 163, 238, 171, 243, 33, 209, 120, 219, 17, 182, 96, 136, 13, 90, 6, 27, 247, 242, 49, 111],
    [179, 64, 123, 119, 226, 139, 161, 127, 36, 251, 218, 88, 20, 217, 212, 85, 112, 85, 185, 193, 230, 181, 4, 22, 54, 219, 135, 98, 235, 180, 182, 7],
    [101, 240, 19, 44, 43, 213, 31, 138, 39, 26, 82, 147, 255, 96, 234, 51, 105, 6, 233, 144, 255, 187, 242, 3, 157, 246, 55, 175, 98, 121, 92, 175],
    [25, 96, 149, 179, 94, 8, 170, 214, 169, 135, 12, 212, 224, 157, 182, 127, 233, 93, 151, 214, 36, 183, 156, 212, 233, 152, 125, 244, 146, 161, 75, 128],
    [247, 43, 130, 141, 234, 172, 61, 187, 109, 31, 56, 30, 14, 232, 92, 158, 48, 161, 108, 234, 170, 180, 233, 77, 200, 248, 45, 152, 125, 11, 1, 171],
]


def _h2(a, b):
    import blake3 as b3
    return b3.blake3(a + b).digest()


def test_merkle_fixtures(oracle):
    l4 = np.array(LEAVES4, dtype=np.uint8)
    n4 = oracle.merkle_nodes(oracle.BLAKE3, l4)
    L = [bytes(r) for r in l4]
    assert n4[1].tobytes() == _h2(_h2(L[0], L[1]), _h2(L[2], L[3]))
    assert not n4[0].any()
    l8 = np.array(LEAVES8, dtype=np.uint8)
    n8 = oracle.merkle_nodes(oracle.BLAKE3, l8)
    L = [bytes(r) for r in l8]
    root = _h2(_h2(_h2(L[0], L[1]), _h2(L[2], L[3])), _h2(_h2(L[4], L[5]), _h2(L[6], L[7])))
    assert n8[1].tobytes() == root
    assert n8[4].tobytes() == _h2(L[0], L[1]) and n8[7].tobytes() == _h2(L[6], L[7])


def _parse_batch_proof(buf):
    depth, nvec = buf[0], buf[1] >> 1   # small vint64 values: one byte, (v<<1)|1
    pos, vecs = 2, []
    for _ in range(nvec):
        ln = buf[pos] >> 1; pos += 1
        vecs.append([buf[pos + 32 * i: pos + 32 * i + 32] for i in range(ln)]); pos += 32 * ln
    assert pos == len(buf)
    return depth, vecs


def test_merkle_prove_batch_fixtures(oracle):
    # merkle/tests.rs:140-186
    l8 = np.array(LEAVES8, dtype=np.uint8)
    n8 = oracle.merkle_nodes(oracle.BLAKE3, l8)
    L = [bytes(r) for r in l8]
    lv, pr = oracle.merkle_prove_batch(l8, n8, [1])
    depth, vecs = _parse_batch_proof(pr)
    assert depth == 3 and lv[0].tobytes() == L[1]
    assert vecs == [[L[0], _h2(L[2], L[3]), _h2(_h2(L[4], L[5]), _h2(L[6], L[7]))]]
    lv, pr = oracle.merkle_prove_batch(l8, n8, [1, 2])
    depth, vecs = _parse_batch_proof(pr)
    assert [x.tobytes() for x in lv] == [L[1], L[2]]
    assert vecs == [[L[0], _h2(_h2(L[4], L[5]), _h2(L[6], L[7]))], [L[3]]]
    lv, pr = oracle.merkle_prove_batch(l8, n8, [1, 6])
    depth, vecs = _parse_batch_proof(pr)
    assert vecs == [[L[0], _h2(L[2], L[3])], [L[7], _h2(L[4], L[5])]]
    lv, pr = oracle.merkle_prove_batch(l8, n8, list(range(8)))
    depth, vecs = _parse_batch_proof(pr)
    assert vecs == [[], [], [], []] and [x.tobytes() for x in lv] == L
    with pytest.raises(ValueError):
        oracle.merkle_prove_batch(l8, n8, [1, 1])


def test_merkle_verify_batch_fixtures(oracle):
    # merkle/tests.rs:189-213 verify_batch: the verifier-side root recomputation (BatchMerkleProof::get_root,
    # proofs.rs:90-205) the oracle verifier uses to accept trace / constraint / FRI openings
    import ctypes as C
    l8 = np.array(LEAVES8, dtype=np.uint8)
    n8 = oracle.merkle_nodes(oracle.BLAKE3, l8)
    root = n8[1].tobytes()
    u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)

    def verify(indexes, lv, pr):
        idx = np.array(indexes, dtype=np.uint64)
        lvb = np.ascontiguousarray(lv, dtype=np.uint8)
        prb = np.frombuffer(pr, dtype=np.uint8)
        rb = np.frombuffer(root, dtype=np.uint8)
        return oracle.lib().wfo_merkle_verify_batch(C.c_int(oracle.BLAKE3), rb.ctypes.data_as(u8p), idx.ctypes.data_as(u64p),
                                                    C.c_size_t(idx.size), lvb.ctypes.data_as(u8p), prb.ctypes.data_as(u8p),
                                                    C.c_size_t(prb.size))

    lv, pr = oracle.merkle_prove_batch(l8, n8, [1])
    assert verify([1], lv, pr) == 0 and verify([2], lv, pr) != 0
    lv, pr = oracle.merkle_prove_batch(l8, n8, [1, 2])
    assert verify([1, 2], lv, pr) == 0
    assert verify([1], lv[:1], pr) != 0 and verify([1, 3], lv, pr) != 0 and verify([1, 2, 3], np.concatenate([lv, lv[:1]]), pr) != 0
    for idx in ([1, 6], [1, 3, 6], list(range(8))):
        lv, pr = oracle.merkle_prove_batch(l8, n8, idx)
        assert verify(idx, lv, pr) == 0
    # a flipped leaf is rejected
    lv, pr = oracle.merkle_prove_batch(l8, n8, [1, 3, 6])
    bad = np.array(lv, dtype=np.uint8).copy()
    bad[1, 0] ^= 1
    assert verify([1, 3, 6], bad, pr) != 0


# ---- FRI ----
def test_transpose_and_fold_positions(oracle):
    # utils/core/src/lib.rs:158-165; fri/src/folding/mod.rs:129-136
    assert list(oracle.transpose_slice(np.arange(8), 2)) == [0, 4, 1, 5, 2, 6, 3, 7]
    assert list(oracle.fold_positions([1, 9, 12, 20], 32, 4)) == [1, 4]


@pytest.mark.parametrize("nf", [2, 4, 8])
@pytest.mark.parametrize("d", [1, 3])
def test_drp_equals_coefficient_folding(oracle, nf, d):
    # fri/src/folding/mod.rs:46-85 (doc-test for N=2; same identity for any N: folded poly has
    # coefficients sum_k alpha^k * poly[N*i + k], evaluated over the domain offset^N * w_{n/N}^i)
    n = 64
    poly = oracle.rand_elems(n // 8 * d, 31)
    poly_full = np.concatenate([poly, np.zeros((n - n // 8) * d, dtype=np.uint64)])
    ev = oracle.evaluate_poly_with_offset(poly_full, 7, 1, d)
    alpha = oracle.rand_elems(d, 32)
    got = oracle.apply_drp(oracle.transpose_slice(ev, nf, d), nf, 7, alpha, d)
    # coefficient-form folding
    m = n // nf
    folded = np.zeros(m * d, dtype=np.uint64)
    for i in range(m):
        acc = np.zeros(d, dtype=np.uint64)
        ak = np.zeros(d, dtype=np.uint64); ak[0] = 1
        for k in range(nf):
            c = poly_full[(nf * i + k) * d:(nf * i + k + 1) * d]
            cd = np.zeros(d, dtype=np.uint64); cd[:] = c
            t = oracle.ext_mul(ak, cd) if d > 1 else np.array([oracle.mul(int(ak[0]), int(cd[0]))], dtype=np.uint64)
            acc = np.array([oracle.add(int(x), int(y)) for x, y in zip(acc, t)], dtype=np.uint64)
            ak = oracle.ext_mul(ak, alpha) if d > 1 else np.array([oracle.mul(int(ak[0]), int(alpha[0]))], dtype=np.uint64)
        folded[i * d:(i + 1) * d] = acc
    want = oracle.evaluate_poly_with_offset(folded, pow(7, nf, P), 1, d)
    assert (got == want).all()


def test_fri_layers_shape(oracle):
    # fri/src/options.rs:85-93 + prover/mod.rs:179-239
    assert oracle.fri_num_layers(2**19, 8, 31, 8) == 4      # cfg1: 2^19 -> 2^16 -> 2^13 -> 2^10 -> 2^7
    n, b = 2**10, 8
    poly = np.concatenate([oracle.rand_elems(n // b, 5), np.zeros(n - n // b, dtype=np.uint64)])
    ev = oracle.evaluate_poly_with_offset(poly, 7, 1)
    roots, rem, alphas = oracle.fri_build_layers(oracle.BLAKE3, ev, 4, 7, b)
    assert roots.shape[0] == 3 and rem.size == 8 and alphas.size == 2   # 1024 -> 256 -> 64 (= 8*8)
    # the remainder is a degree < 8 polynomial: re-evaluating it over the last domain reproduces
    # the twice-folded codeword
    t = ev
    for a in alphas:
        t = oracle.apply_drp(oracle.transpose_slice(t, 4), 4, 7, [a])
    back = oracle.interpolate_poly_with_offset(t, 7)
    assert (back[:8][::-1] == rem).all() and not back[8:].any()


def test_random_coin(oracle):
    import blake3 as b3
    c = oracle.RandomCoin(oracle.BLAKE3, [1, 2, 3])
    seed = b3.blake3(np.array([1, 2, 3], dtype="<u8").tobytes()).digest()
    assert c.seed == seed
    # draw: merge_with_int(seed, ++counter), first 8 bytes LE, rejected if >= p (default.rs:156-170)
    cnt, want = 0, None
    while want is None:
        cnt += 1
        v = int.from_bytes(b3.blake3(seed + cnt.to_bytes(8, "little")).digest()[:8], "little")
        if v < P:
            want = v
    assert int(c.draw(1)[0]) == want
    c.reseed(bytes(32))
    assert c.seed == b3.blake3(seed + bytes(32)).digest()
    lz = c.leading_zeros(5)
    head = int.from_bytes(b3.blake3(c.seed + (5).to_bytes(8, "little")).digest()[:8], "little")
    assert lz == (head & -head).bit_length() - 1
    ints = c.draw_integers(20, 1024, 77)
    assert len(ints) == 20 and all(int(v) < 1024 for v in ints)


def test_context_and_options_to_elements(oracle):
    # air/src/proof/context.rs tests::context_to_elements, air/src/air/trace_info.rs tests::trace_info_to_elements,
    # air/src/options.rs tests::proof_options_to_elements: exact packing of the channel seed prefix
    import ctypes as C
    L = oracle.lib()
    L.wfo_context_elements.restype = C.c_size_t
    out = np.zeros(16, dtype=np.uint64)
    opts = oracle.make_opts(num_queries=30, blowup=8, grinding=20, ext=1, folding=8, rem_max_deg=127)
    n = L.wfo_context_elements(C.c_size_t(20), C.c_size_t(9), C.c_size_t(12), C.c_size_t(4096), C.c_size_t(128),
                               opts.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    first = int.from_bytes(bytes([12, 9, 1, 20]), "little")            # [aux_rands, aux_width, num_aux_segments, main_width]
    ext_fri = int.from_bytes(bytes([8, 127, 8, 1]), "little")          # [blowup, remainder degree, folding, FieldExtension::None = 1]
    assert [int(x) for x in out[:n]] == [first, 4096, 1, 0xFFFFFFFF, 128, ext_fri, 20, 30]
    # single-segment trace info: [num_aux_segments = 0, main_width, 0, 0]
    n = L.wfo_context_elements(C.c_size_t(20), C.c_size_t(0), C.c_size_t(0), C.c_size_t(64), C.c_size_t(5),
                               opts.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert [int(x) for x in out[:2]] == [int.from_bytes(bytes([0, 20, 0, 0]), "little"), 64]


def test_partition_sizes(oracle):
    # air/src/options.rs tests::correct_partition_sizes
    import ctypes as C
    L = oracle.lib()
    L.wfo_partition_size.restype = C.c_size_t
    L.wfo_num_partitions.restype = C.c_size_t
    ps = lambda np_, rate, d, cols: L.wfo_partition_size(C.c_size_t(np_), C.c_size_t(rate), C.c_size_t(d), C.c_size_t(cols))
    npn = lambda np_, rate, d, cols: L.wfo_num_partitions(C.c_size_t(np_), C.c_size_t(rate), C.c_size_t(d), C.c_size_t(cols))
    assert (ps(4, 8, 1, 7), npn(4, 8, 1, 7)) == (8, 1)
    assert (ps(4, 8, 1, 70), npn(4, 8, 1, 70)) == (18, 4)
    assert (ps(2, 8, 3, 7), npn(2, 8, 3, 7)) == (4, 2)
    assert (ps(4, 8, 3, 7), npn(4, 8, 3, 7)) == (2, 4)
    assert (ps(4, 8, 3, 3), npn(4, 8, 3, 3)) == (2, 2)


def test_boundary_constraint_groups_kat(oracle):
    # air/src/air/tests.rs::get_boundary_constraints (:64-230): eight assertions on a 16-step trace ->
    # five groups ordered by (stride, first step), divisor x^a - g^(a * first_step), coefficients handed out
    # in the sorted-assertion order, sequence polynomials = interpolants over the size-len subgroup with
    # x offset g^(-first_step). Expected values are recomputed here with Python integers.
    import ctypes as C
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import airs
    P = airs.P
    A = airs.AirBuilder(2)
    A.constraint(A.sub(A.nxt(0), A.cur(0)), 1)
    A.assert_single(0, 0, 3)
    A.assert_single(0, 9, 5)
    A.assert_single(1, 9, 9)
    A.assert_sequence(0, 2, 4, [1, 2, 3, 4])
    A.assert_sequence(1, 2, 4, [1, 2, 3, 4])
    A.assert_sequence(1, 0, 8, [1, 2])
    A.assert_sequence(0, 3, 8, [1, 2])
    A.assert_periodic(1, 3, 8, 7)
    desc = A.build()
    n = 16
    g = oracle.root_of_unity(4)
    coeffs = np.array([1000 + i for i in range(8)], dtype=np.uint64)   # draw order = sorted-assertion order
    out = np.zeros(256, dtype=np.uint64)
    L = oracle.lib()
    L.wfo_boundary_groups.restype = C.c_long
    u64p = C.POINTER(C.c_uint64)
    k = L.wfo_boundary_groups(desc.ctypes.data_as(u64p), C.c_size_t(desc.size), C.c_size_t(n), coeffs.ctypes.data_as(u64p),
                              out.ctypes.data_as(u64p), C.c_size_t(out.size))
    assert k > 0
    w = [int(x) for x in out[:k]]
    pos = 1
    groups = []
    for _ in range(w[0]):
        a, b, ne = w[pos:pos + 3]
        pos += 3
        ents = []
        for _ in range(ne):
            col, cc, xo, pl = w[pos:pos + 4]
            pos += 4
            ents.append((col, cc, xo, w[pos:pos + pl]))
            pos += pl
        groups.append((a, b, ents))
    assert pos == k

    def interp(values):  # polynom::interpolate over the subgroup of size len(values) (tests.rs:305-311)
        m = len(values)
        h = pow(g, n // m, P)
        xs = [pow(h, i, P) for i in range(m)]
        coef = [0] * m
        for i, (xi, yi) in enumerate(zip(xs, values)):  # Lagrange basis, expanded
            num, den = [1], 1
            for j, xj in enumerate(xs):
                if j != i:
                    num = [(a0 - xj * a1) % P for a0, a1 in zip([0] + num, num + [0])]
                    den = den * (xi - xj) % P
            sc = yi * pow(den, P - 2, P) % P
            coef = [(c + sc * t) % P for c, t in zip(coef, num)]
        return coef

    ginv = pow(g, P - 2, P)
    want = [
        (1, pow(g, 0, P), [(0, 1000, 1, [3])]),                                                   # group 0: step 0
        (1, pow(g, 9, P), [(0, 1001, 1, [5]), (1, 1002, 1, [9])]),                                # group 1: step 9
        (4, pow(g, 8, P), [(0, 1003, pow(ginv, 2, P), interp([1, 2, 3, 4])), (1, 1004, pow(ginv, 2, P), interp([1, 2, 3, 4]))]),
        (2, pow(g, 0, P), [(1, 1005, 1, interp([1, 2]))]),                                        # group 3: steps 0, 8
        (2, pow(g, 6, P), [(0, 1006, pow(ginv, 3, P), interp([1, 2])), (1, 1007, 1, [7])]),       # group 4: steps 3, 11
    ]
    assert groups == want
    # air/src/air/divisor.rs::constraint_divisor_equivalence: x^a - b vanishes exactly on the asserted steps
    asserted = [{0}, {9}, {2, 6, 10, 14}, {0, 8}, {3, 11}]
    for (a_, b_, _), steps in zip(groups, asserted):
        zeros = {i for i in range(n) if (pow(pow(g, i, P), a_, P) - b_) % P == 0}
        assert zeros == steps


def test_usize_vint64_encoding(oracle):
    # utils/core/src/tests.rs::write_serializable_usize: encoded lengths 1, 1, 2, 3, 9 and round trip
    import ctypes as C
    L = oracle.lib()
    L.wfo_write_usize.restype = C.c_size_t
    buf = (C.c_uint8 * 9)()
    total = 0
    for v, want_total in [(0, 1), (1, 2), (255, 4), (234567, 7), (2**64 - 1, 16)]:
        ln = L.wfo_write_usize(C.c_uint64(v), buf)
        total += ln
        assert total == want_total
        b = bytes(buf[:ln])
        # reader (byte_reader.rs read_usize): length from the trailing zeros of the first byte
        first = b[0]
        length = 9 if first == 0 else ((first & -first).bit_length())
        assert length == ln
        got = int.from_bytes(b[1:9], "little") if length == 9 else int.from_bytes(b, "little") >> length
        assert got == v


def test_periodic_value_table(oracle):
    # prover/src/constraints/evaluator/periodic_table.rs::periodic_value_table: trace length 32, columns [1, 2] and
    # [3, 4, 5, 6]; row i of the CE domain = poly_j((7 w_ce^i)^(n / L_j)), polys = interpolants over the cycle
    import ctypes as C
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import airs
    P = airs.P
    n = 32
    A = airs.AirBuilder(1)
    A.periodic = [[1, 2], [3, 4, 5, 6]]
    A.constraint(A.mul(A.mul(A.cur(0), A.per(0)), A.per(1)), 1, [2, 4])   # any constraint that gives ce_blowup 4
    A.assert_single(0, 0, 0)
    desc = A.build()
    L = oracle.lib()
    u64p = C.POINTER(C.c_uint64)
    L.wfo_ce_blowup.restype = C.c_size_t
    ceb = L.wfo_ce_blowup(desc.ctypes.data_as(u64p), C.c_size_t(desc.size))
    ce = n * ceb
    g_ce = oracle.root_of_unity(ce.bit_length() - 1)

    def interp_eval(values, x):  # Lagrange over the subgroup of size len(values)
        m = len(values)
        h = pow(oracle.root_of_unity(5), n // m, P)
        xs = [pow(h, i, P) for i in range(m)]
        acc = 0
        for i, (xi, yi) in enumerate(zip(xs, values)):
            num, den = 1, 1
            for j, xj in enumerate(xs):
                if j != i:
                    num = num * (x - xj) % P
                    den = den * (xi - xj) % P
            acc = (acc + yi * num * pow(den, P - 2, P)) % P
        return acc

    row = np.zeros(2, dtype=np.uint64)
    for i in range(ce):
        assert L.wfo_periodic_row(desc.ctypes.data_as(u64p), C.c_size_t(desc.size), C.c_size_t(n), C.c_size_t(i), row.ctypes.data_as(u64p)) == 2
        x = 7 * pow(g_ce, i, P) % P
        assert int(row[0]) == interp_eval([1, 2], pow(x, n // 2, P))
        assert int(row[1]) == interp_eval([3, 4, 5, 6], pow(x, n // 4, P))


def test_split_radix_fft_equals_serial_network(oracle):
    o = oracle
    # math/src/fft/tests.rs checks the concurrent transforms against the serial ones; the oracle's restatement of
    # concurrent::split_radix_fft (the CPU baseline's parallel decomposition) must give the serial network's words
    for log_n in (10, 11, 12, 13):           # even sizes: stretch 1, odd sizes: stretch 2
        for d in (1, 3):
            v = o.rand_elems(((1 << log_n) * d,), 77 + log_n + d)
            for inverse in (False, True):
                want = o.fft_in_place(v, d, inverse)
                for nt in (1, 3, 8):
                    assert (o.fft_in_place(v, d, inverse, split_radix_threads=nt) == want).all(), (log_n, d, inverse, nt)
