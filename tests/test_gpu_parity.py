"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs — bit-exact (integer arithmetic mod p, bytes of digests)."""
import numpy as np
import pytest

import winterfell_b200 as wf

pytestmark = pytest.mark.gpu
P = wf.P


@pytest.fixture(scope="module")
def ctx():
    c = wf.Context(0)
    yield c
    c.close()


def naive_eval(p, x):
    acc = 0
    for c in reversed([int(v) for v in p]):
        acc = (acc * x + c) % P
    return acc


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 8, 10, 11, 12, 13, 16])
@pytest.mark.parametrize("cols", [1, 2, 3, 8, 11])
def test_ntt_intt_vs_oracle(ctx, oracle, log_n, cols):
    n = 1 << log_n
    x = oracle.rand_elems((cols, n), 1000 * log_n + cols)
    m = ctx.mat_from_host_columns(x)
    assert (m.to_columns() == x).all() and (m.to_rows() == x.T).all()
    ev = m.evaluate()
    got = ev.to_columns()
    for c in range(cols):
        assert (got[c] == oracle.evaluate_poly(x[c])).all(), (log_n, c)
    back = ev.interpolate()
    assert (back.to_columns() == x).all()
    it = m.interpolate().to_columns()
    assert (it[0] == oracle.interpolate_poly(x[0])).all()
    for h in (m, ev, back):
        h.free()


def test_ntt_small_vs_naive(ctx, oracle):
    # math/src/fft/tests.rs:20-61 on the device path
    for n in (4, 8, 16):
        p = oracle.rand_elems((1, n), n)
        ev = ctx.mat_from_host_columns(p).evaluate().to_columns()[0]
        g = oracle.root_of_unity(n.bit_length() - 1)
        for i in range(n):
            assert int(ev[i]) == naive_eval(p[0], pow(g, i, P))


def test_montgomery_abi(ctx, oracle):
    # host buffers in the reference's in-memory form (f64/mod.rs:57-64): x * 2^64 mod p
    x = oracle.rand_elems((3, 64), 9)
    xm = np.array([[oracle.to_mont(int(v)) for v in row] for row in x], dtype=np.uint64)
    m = ctx.mat_from_host_columns(xm, mont=True)
    assert (m.to_columns() == x).all()
    assert (m.to_columns(mont=True) == xm).all()
    assert (m.read_rows([5, 0, 63], mont=True) == xm.T[[5, 0, 63]]).all()


@pytest.mark.parametrize("log_n,cols,log_b", [(3, 2, 3), (8, 64, 3), (10, 8, 3), (11, 5, 2), (12, 8, 3), (13, 3, 1), (14, 16, 3), (16, 2, 3)])
def test_lde_vs_oracle(ctx, oracle, log_n, cols, log_b):
    # prover/src/matrix/tests.rs:15-42 (64 cols, n=256, blowup 8) and other shapes
    n = 1 << log_n
    polys = oracle.rand_elems((cols, n), 77 + log_n)
    want = oracle.lde_rows(polys, 1 << log_b)
    m = ctx.mat_from_host_columns(polys)
    lde = m.lde(log_b)
    assert (lde.to_rows() == want).all()
    pos = [0, 1, (n << log_b) - 1, 12345 % (n << log_b)]
    assert (lde.read_rows(pos) == want[pos]).all()
    m.free(); lde.free()


@pytest.mark.parametrize("d", [2, 3])
def test_lde_extension_columns(ctx, oracle, d):
    # ColMatrix<E> with E = quadratic / cubic extension: d base columns per column
    n, c, b = 256, 3, 8
    polys = oracle.rand_elems((c, n * d), 5 + d)
    want = oracle.lde_rows(polys, b, d)
    m = ctx.mat_from_host_columns(polys, ext_degree=d)
    assert m.cols == c * d
    assert (m.lde(3).to_rows() == want).all()


def test_interpolate_with_offset(ctx, oracle):
    n = 1 << 12
    ev = oracle.rand_elems((2, n), 4)
    got = ctx.mat_from_host_columns(ev).interpolate_with_offset(7).to_columns()
    for c in range(2):
        assert (got[c] == oracle.interpolate_poly_with_offset(ev[c], 7)).all()


@pytest.mark.parametrize("h", [wf.HASH_BLAKE3_256, wf.HASH_RP64_256, wf.HASH_RPJIVE64_256, wf.HASH_BLAKE3_192, wf.HASH_SHA3_256])
@pytest.mark.parametrize("cols", [1, 2, 3, 4, 7, 8, 9, 16, 24, 64, 128, 130])
def test_row_hash_and_merkle_vs_oracle(ctx, oracle, h, cols):
    rows = 1024 if h in (wf.HASH_BLAKE3_256, wf.HASH_BLAKE3_192) else 256
    x = oracle.rand_elems((cols, rows), 31 * cols + h)
    m = ctx.mat_from_host_columns(x)
    t = ctx.commit_rows(h, m)
    lv, nd = t.to_host()
    want_lv = oracle.hash_rows(h, np.ascontiguousarray(x.T))
    assert (lv == want_lv).all()
    want_nd = oracle.merkle_nodes(h, want_lv)
    assert (nd == want_nd).all()
    assert t.root() == want_nd[1].tobytes()
    # batch openings (crypto/src/merkle/mod.rs:217-272)
    for pos in ([1], [1, 2], [1, 6], [3, 4, 5], [0, 7, 100, 101, rows - 1], list(range(8))):
        glv, gpr = t.open_many(pos)
        wlv, wpr = oracle.merkle_prove_batch(want_lv, want_nd, pos, h)
        assert (glv == wlv).all() and gpr == wpr
    m.free(); t.free()


@pytest.mark.parametrize("nleaves", [2, 4, 8, 256, 512, 1024, 4096])
def test_merkle_small_trees(ctx, oracle, nleaves):
    rng = np.random.default_rng(nleaves)
    leaves = rng.integers(0, 256, size=(nleaves, 32), dtype=np.uint8)
    t = ctx.tree_from_leaves(wf.HASH_BLAKE3_256, leaves)
    lv, nd = t.to_host()
    assert (nd == oracle.merkle_nodes(oracle.BLAKE3, leaves)).all()


def test_merkle_reference_fixture(ctx):
    # crypto/src/merkle/tests.rs:14-84 (LEAVES8): root == nested hash_2x1
    import blake3 as b3
    from test_oracle_kats import LEAVES8
    l8 = np.array(LEAVES8, dtype=np.uint8)
    L = [bytes(r) for r in l8]
    h2 = lambda a, b: b3.blake3(a + b).digest()
    root = h2(h2(h2(L[0], L[1]), h2(L[2], L[3])), h2(h2(L[4], L[5]), h2(L[6], L[7])))
    assert ctx.tree_from_leaves(wf.HASH_BLAKE3_256, l8).root() == root


@pytest.mark.parametrize("h", [wf.HASH_BLAKE3_256, wf.HASH_RP64_256, wf.HASH_RPJIVE64_256, wf.HASH_BLAKE3_192, wf.HASH_SHA3_256])
@pytest.mark.parametrize("d,nf,log_len", [(1, 4, 12), (1, 2, 10), (1, 8, 12), (1, 16, 12), (2, 4, 10), (3, 4, 12), (3, 8, 9)])
def test_fri_layers_vs_oracle(ctx, oracle, h, d, nf, log_len):
    # fri/src/prover/tests.rs round trip shape: commit phase roots and remainder
    L, b = 1 << log_len, 8
    poly = np.concatenate([oracle.rand_elems(L // b * d, 3 + d), np.zeros((L - L // b) * d, dtype=np.uint64)])
    ev = oracle.evaluate_poly_with_offset(poly, 7, 1, d)
    want_roots, want_rem, alphas = oracle.fri_build_layers(h, ev, nf, 7, b, d)
    cols = np.ascontiguousarray(ev.reshape(L, d).T)     # d base columns
    m = ctx.mat_from_host_columns(cols)
    f, roots = ctx.fri_build_layers_default(h, m, d, nf, 7, b)
    assert (roots == want_roots).all()
    assert (f.remainder() == want_rem).all()
    f.free(); m.free()


def test_fri_fold_dev_vs_oracle(ctx, oracle):
    import torch
    for d, nf in ((1, 4), (3, 4), (1, 8), (2, 2)):
        L = 1 << 12
        ev = oracle.rand_elems(L * d, 11 * d + nf)
        alpha = oracle.rand_elems(d, 5)
        want = oracle.apply_drp(oracle.transpose_slice(ev, nf, d), nf, 7, alpha, d)
        t_in = torch.from_numpy(ev.view(np.int64)).cuda()
        t_out = torch.empty(L // nf * d, dtype=torch.int64, device="cuda")
        ctx.fri_fold_dev(t_in.data_ptr(), L, d, nf, alpha, t_out.data_ptr())
        ctx.sync()
        assert (t_out.cpu().numpy().view(np.uint64) == want).all()


def test_plain_dev_kernels(ctx, oracle):
    import torch
    n, c = 1 << 13, 5
    x = oracle.rand_elems((c, n), 8)
    t = torch.from_numpy(x.view(np.int64)).cuda()
    ctx.ntt_dev(t.data_ptr(), 13, c, False)
    ctx.sync()
    got = t.cpu().numpy().view(np.uint64)
    for j in range(c):
        assert (got[j] == oracle.evaluate_poly(x[j])).all()
    ctx.ntt_dev(t.data_ptr(), 13, c, True)
    ctx.sync()
    assert (t.cpu().numpy().view(np.uint64) == x).all()
    rows = np.ascontiguousarray(x.T)
    tr = torch.from_numpy(rows.view(np.int64)).cuda()
    dg = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    nd = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    ctx.hash_rows_dev(wf.HASH_BLAKE3_256, tr.data_ptr(), n, c, dg.data_ptr())
    ctx.merkle_dev(wf.HASH_BLAKE3_256, dg.data_ptr(), n, nd.data_ptr())
    ctx.sync()
    want = oracle.hash_rows(oracle.BLAKE3, rows)
    assert (dg.cpu().numpy().reshape(n, 32) == want).all()
    assert (nd.cpu().numpy().reshape(n, 32) == oracle.merkle_nodes(oracle.BLAKE3, want)).all()


@pytest.mark.parametrize("h", [wf.HASH_BLAKE3_256, wf.HASH_RP64_256, wf.HASH_RPJIVE64_256, wf.HASH_BLAKE3_192, wf.HASH_SHA3_256])
@pytest.mark.parametrize("cols,psize", [(64, 8), (64, 16), (20, 8), (9, 4), (130, 9), (16, 1)])
def test_partitioned_row_hash_vs_oracle(ctx, oracle, h, cols, psize):
    # RowMatrix::commit_to_rows with PartitionOptions (row_matrix.rs:204-223): merge_many of chunk digests
    rows = 128
    x = oracle.rand_elems((cols, rows), cols * 7 + psize)
    m = ctx.mat_from_host_columns(x)
    t = ctx.commit_rows(h, m, partition_size=psize)
    lv, nd = t.to_host()
    want = oracle.hash_rows(h, np.ascontiguousarray(x.T), partition_size=psize)
    assert (lv == want).all()
    assert (nd == oracle.merkle_nodes(h, want)).all()


def test_field_ops_edge_cases(ctx):
    """Device field arithmetic (gl64.cuh) on operands chosen to reach every reduction path. Uniform random
    data takes the canonicalisation step (result in [p, 2^64) before the final subtraction) with
    probability 2^-32 per operation, so the NTT / proof parity tests never see it: products that land in
    [0, 2^32) and operands next to 0, 2^32, p are generated here and checked against Python integers."""
    import torch
    P = wf.P
    rng = np.random.default_rng(99)
    edge = [0, 1, 2, 3, P - 1, P - 2, P - 3, 2**32 - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**33 - 1, 2**63, 2**63 - 1,
            P - 2**32, P - 2**32 + 1, P - 2**32 - 1, 0xFFFFFFFF00000000, 0xFFFFFFFE00000001, 0x00000001FFFFFFFF, 2**48, 2**24]
    a, b = [], []
    for x in edge:
        for y in edge:
            a.append(x); b.append(y)
    small = [0, 1, 2, 5, 2**16, 2**31, 2**32 - 3, 2**32 - 2, 2**32 - 1, 2**32, 2**32 + 1, P - 1, P - 2, P - 2**32, P - 2**32 + 5]
    for _ in range(6000):  # a * b = c with c in the ranges where the pre-canonical result exceeds p
        x = int(rng.integers(1, 2**63)) * 2 % P or 1
        c = small[int(rng.integers(0, len(small)))] if rng.random() < 0.5 else int(rng.integers(0, 2**32))
        a.append(x); b.append(c * pow(x, P - 2, P) % P)
    for _ in range(4000):
        a.append(int(rng.integers(0, 2**63)) * 2 % P); b.append(int(rng.integers(0, 2**63)) * 2 % P)
    for _ in range(2000):  # sums / differences around the wrap points
        x = int(rng.integers(0, 2**63)) * 2 % P
        d = int(rng.integers(0, 4)) - 2
        a.append(x); b.append((P - x + d) % P)
    n = len(a)
    ta = torch.from_numpy(np.array(a, dtype=np.uint64).view(np.int64)).cuda()
    tb = torch.from_numpy(np.array(b, dtype=np.uint64).view(np.int64)).cuda()
    out = torch.empty(22 * n, dtype=torch.int64, device="cuda")
    ctx.field_ops_dev(ta.data_ptr(), tb.data_ptr(), n, out.data_ptr())
    ctx.sync()
    got = out.cpu().numpy().view(np.uint64).reshape(22, n)
    shifts = [1, 3, 6, 12, 24, 31, 32, 33, 48, 63, 64, 65, 72, 80, 84, 90, 95, 96]
    for i in range(n):
        x, y = a[i], b[i]
        assert int(got[0, i]) == x * y % P, ("mul", hex(x), hex(y))
        assert int(got[1, i]) == (x + y) % P, ("add/butterfly", hex(x), hex(y))
        assert int(got[2, i]) == (x - y) % P, ("sub", hex(x), hex(y))
        assert int(got[3, i]) == (pow(x, P - 2, P) if x else 0), ("inv", hex(x))
        for k, sh in enumerate(shifts):
            assert int(got[4 + k, i]) == (x << sh) % P, ("shift", sh, hex(x))


def _py_ext_mul(d, a, b):
    P = wf.P
    if d == 2:   # x^2 = x - 2 (math/src/field/f64/mod.rs:403-409)
        return [(a[0] * b[0] - 2 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0] + a[1] * b[1]) % P]
    # x^3 = x + 1 (mod.rs:445-466)
    c = [0] * 5
    for i in range(3):
        for j in range(3):
            c[i + j] += a[i] * b[j]
    c[2] += c[4]; c[1] += c[4]      # x^4 = x^2 + x
    c[1] += c[3]; c[0] += c[3]      # x^3 = x + 1
    return [c[0] % P, c[1] % P, c[2] % P]


def _py_ext_pow(d, a, e):
    r = [1] + [0] * (d - 1)
    while e:
        if e & 1:
            r = _py_ext_mul(d, r, a)
        a = _py_ext_mul(d, a, a)
        e >>= 1
    return r


@pytest.mark.parametrize("d", [2, 3])
def test_extension_field_ops_on_reference_vectors(ctx, d):
    """Device quadratic / cubic extension arithmetic (gl64.cuh ext_mul / ext_inv / ext_frobenius / ext_mul_base) on the
    reference's own product vectors, including its two overflow cases per extension (math/src/field/f64/tests.rs:220-246,
    288-345), its conjugate vectors (:259-283), and on edge / random operands against big-integer arithmetic in the
    quotient ring; inverses are checked by a * a^-1 = 1 and against a^(p^d - 2)."""
    import torch
    P = wf.P
    m = P
    if d == 2:
        kat = [([3, 1], [4, 2], [8, 12]), ([3, m - 1], [m - 3, 5], [1, 13]), ([3, m - 1], [10, m - 2], [26, 18446744069414584307])]
        conj = [([m - 1, 3], [2, 18446744069414584318]), ([m - 3, m - 2], [18446744069414584316, 2]), ([4, 7], [11, 18446744069414584314])]
    else:
        kat = [([3, 5, 2], [320, 68, 3], [1111, 1961, 995]),
               ([18446744069414584267, 18446744069414584309, 9223372034707292160], [18446744069414584101, 420, 18446744069414584121],
                [14070, 18446744069414566571, 5970]),
               ([18446744069414584266, 18446744069412558094, 5268562], [18446744069414583589, 1226, 5346],
                [18446744065041672051, 25275910656, 21824696736])]
        conj = []
    for a, b, want in kat:   # the test's own model must reproduce the reference vectors first
        assert _py_ext_mul(d, a, b) == want
    rng = np.random.default_rng(7 + d)
    edge = [0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, P - 2**32, 2**63]
    A = [k[0] for k in kat] + [c[0] for c in conj]
    B = [k[1] for k in kat] + [[1] + [0] * (d - 1) for _ in conj]
    for _ in range(300):
        A.append([edge[int(rng.integers(0, len(edge)))] for _ in range(d)])
        B.append([edge[int(rng.integers(0, len(edge)))] for _ in range(d)])
    for _ in range(1500):
        A.append([int(rng.integers(0, 2**63)) * 2 % P for _ in range(d)])
        B.append([int(rng.integers(0, 2**63)) * 2 % P for _ in range(d)])
    n = len(A)
    ta = torch.from_numpy(np.array(A, dtype=np.uint64).view(np.int64)).cuda()
    tb = torch.from_numpy(np.array(B, dtype=np.uint64).view(np.int64)).cuda()
    out = torch.empty(6 * n * d, dtype=torch.int64, device="cuda")
    ctx.ext_ops_dev(d, ta.data_ptr(), tb.data_ptr(), n, out.data_ptr())
    ctx.sync()
    got = out.cpu().numpy().view(np.uint64).reshape(6, n, d)
    for i, (a, b, want) in enumerate(kat):
        assert [int(v) for v in got[0, i]] == want
    for i, (a, want) in enumerate(conj):
        assert [int(v) for v in got[2, len(kat) + i]] == want     # Frobenius = conjugation in the quadratic extension
    one = [1] + [0] * (d - 1)
    for i in range(n):
        a, b = A[i], B[i]
        assert [int(v) for v in got[0, i]] == _py_ext_mul(d, a, b), ("mul", a, b)
        inv = [int(v) for v in got[1, i]]
        if any(a):
            assert _py_ext_mul(d, a, inv) == one, ("inv", a)
        else:
            assert inv == [0] * d
        assert [int(v) for v in got[2, i]] == _py_ext_pow(d, a, P), ("frobenius", a)
        assert [int(v) for v in got[3, i]] == [x * b[0] % P for x in a]
        assert [int(v) for v in got[4, i]] == [(x + y) % P for x, y in zip(a, b)]
        assert [int(v) for v in got[5, i]] == [(x - y) % P for x, y in zip(a, b)]
    for i in range(0, 40):  # a^-1 == a^(p^d - 2) on a sample (slow in Python)
        if any(A[i]):
            assert [int(v) for v in got[1, i]] == _py_ext_pow(d, A[i], P**d - 2)


@pytest.mark.parametrize("ncols,log_n", [(8, 13), (16, 12), (12, 13), (3, 12), (5, 14), (8, 9), (24, 12)])
def test_pipelined_trace_lde_equals_stepwise(ctx, oracle, ncols, log_n):
    # wf_trace_lde_from_host cuts the columns into chunks (whole segments for >= 2 segments, halves of the
    # one segment otherwise; small or narrow inputs take the plain path) and overlaps upload with compute:
    # coefficients and LDE must equal from_host_columns -> interpolate -> lde, and the oracle
    cols = oracle.rand_elems((ncols, 1 << log_n), 1000 + ncols)
    polys, lde = ctx.trace_lde_from_host(cols, 3)
    m = ctx.mat_from_host_columns(cols)
    p2 = m.interpolate()
    l2 = p2.lde(3)
    assert (polys.to_columns() == p2.to_columns()).all()
    assert (lde.to_rows() == l2.to_rows()).all()
    if log_n <= 12:
        assert (lde.to_rows() == oracle.lde_rows(oracle.interpolate_columns(cols), 8)).all()
    to_m = np.vectorize(lambda v: oracle.to_mont(int(v)), otypes=[np.uint64])
    if ncols in (3, 8, 16) and log_n <= 13:  # Montgomery-form input through the plain, half-segment and whole-segment paths
        pm, lm = ctx.trace_lde_from_host(to_m(cols), 3, mont=True)
        assert (lm.to_rows() == l2.to_rows()).all()
        pm.free(); lm.free()
    tree_a, tree_b = ctx.commit_rows(wf.HASH_BLAKE3_256, lde), ctx.commit_rows(wf.HASH_BLAKE3_256, l2)
    assert tree_a.root() == tree_b.root()
    for o in (polys, lde, m, p2, l2, tree_a, tree_b):
        o.free()
