"""AIR descriptions for the generic prover (flat u64 format shared by oracle/wf_prover.cpp parse_air and
the product's wf_prove_air) and matching trace builders. Each AIR restates a reference example over
the f64 field, or exercises a feature (periodic columns, periodic assertions, degree > 1)."""
import numpy as np

P = 0xFFFFFFFF00000001
ADD, SUB, MUL, CONST, OUT = 0, 1, 2, 3, 4


class AirBuilder:
    def __init__(self, width):
        self.w = width
        self.degrees, self.periodic, self.consts, self.prog, self.asserts, self.pub = [], [], [], [], [], []
        self.exemptions = 1
        self.next_reg = None

    def cur(self, c): return c
    def nxt(self, c): return self.w + c
    def per(self, j): return 2 * self.w + j

    def _tmp(self):
        if self.next_reg is None:
            self.next_reg = 2 * self.w + len(self.periodic)
        r = self.next_reg
        self.next_reg += 1
        return r

    def op(self, code, a, b):
        d = self._tmp()
        self.prog.append((code, d, a, b))
        return d

    def add(self, a, b): return self.op(ADD, a, b)
    def sub(self, a, b): return self.op(SUB, a, b)
    def mul(self, a, b): return self.op(MUL, a, b)

    def const(self, v):
        self.consts.append(v % P)
        d = self._tmp()
        self.prog.append((CONST, d, len(self.consts) - 1, 0))
        return d

    def constraint(self, reg, base_degree, cycles=()):
        self.prog.append((OUT, len(self.degrees), reg, 0))
        self.degrees.append((base_degree, list(cycles)))

    # Assertion::single / ::periodic / ::sequence (air/src/air/assertions/mod.rs:62-120)
    def assert_single(self, column, step, value): self.asserts.append((column, step, 0, 1, int(value) % P))
    def assert_periodic(self, column, first_step, stride, value): self.asserts.append((column, first_step, stride, 1, int(value) % P))

    def assert_sequence(self, column, first_step, stride, values):
        self.asserts.append((column, first_step, stride, len(values)) + tuple(int(v) % P for v in values))

    def aux(self, aux_width, num_rands):
        """Declares the auxiliary segment; returns the builder for its constraint program."""
        self.aux_seg = AuxSegment(self, aux_width, num_rands)
        return self.aux_seg

    def build(self):
        d = self._build_main()
        seg = getattr(self, "aux_seg", None)
        if seg is not None:
            d = np.concatenate([d, seg.build()])
        return d

    def _build_main(self):
        if self.next_reg is None:
            self.next_reg = 2 * self.w + len(self.periodic)
        d = [self.w, len(self.degrees)]
        for base, cyc in self.degrees:
            d += [base, len(cyc)] + list(cyc)
        d.append(len(self.periodic))
        for col in self.periodic:
            d += [len(col)] + [int(v) % P for v in col]
        d += [len(self.consts)] + self.consts
        d += [self.next_reg, len(self.prog)]
        for ins in self.prog:
            d += list(ins)
        d.append(len(self.asserts))
        for a in self.asserts:
            d += list(a)
        d += [len(self.pub)] + [int(v) % P for v in self.pub]
        d.append(self.exemptions)
        return np.array(d, dtype=np.uint64)


class AuxSegment:
    """Aux-segment constraint program (Air::evaluate_aux_transition, air/src/air/mod.rs:248-260).
    Registers: main cur/next, aux cur/next, periodic values, random elements, temporaries — all in E."""

    def __init__(self, air, aw, nr):
        self.air, self.aw, self.nr = air, aw, nr
        self.degrees, self.prog, self.asserts = [], [], []
        self.next_reg = 2 * air.w + 2 * aw + len(air.periodic) + nr

    def cur(self, c): return c
    def nxt(self, c): return self.air.w + c
    def acur(self, c): return 2 * self.air.w + c
    def anxt(self, c): return 2 * self.air.w + self.aw + c
    def per(self, j): return 2 * self.air.w + 2 * self.aw + j
    def rnd(self, j): return 2 * self.air.w + 2 * self.aw + len(self.air.periodic) + j

    def op(self, code, a, b):
        d = self.next_reg
        self.next_reg += 1
        self.prog.append((code, d, a, b))
        return d

    def add(self, a, b): return self.op(ADD, a, b)
    def sub(self, a, b): return self.op(SUB, a, b)
    def mul(self, a, b): return self.op(MUL, a, b)

    def const(self, v):
        self.air.consts.append(v % P)
        return self.op(CONST, len(self.air.consts) - 1, 0)

    def constraint(self, reg, base_degree, cycles=()):
        self.prog.append((OUT, len(self.degrees), reg, 0))
        self.degrees.append((base_degree, list(cycles)))

    def assert_single(self, column, step, value=(0, 0, 0)):
        self.asserts.append((column, step, 0, 1) + tuple(int(v) % P for v in value))

    def assert_sequence(self, column, first_step, stride, values):
        """values: list of (v0, v1, v2) elements of E, one per asserted step."""
        flat = tuple(int(x) % P for v in values for x in v)
        self.asserts.append((column, first_step, stride, len(values)) + flat)

    def build(self):
        d = [self.aw, self.nr, len(self.degrees)]
        for base, cyc in self.degrees:
            d += [base, len(cyc)] + list(cyc)
        d += [self.next_reg, len(self.prog)]
        for ins in self.prog:
            d += list(ins)
        d.append(len(self.asserts))
        for a in self.asserts:
            d += list(a)
        return np.array(d, dtype=np.uint64)


def fib_small_x(k, n):
    """k copies of examples/src/fibonacci/fib_small (pair j starts at (j+1, j+1)); must produce exactly
    the proofs of the specialised wf_prove_fib / wfo_prove_fib path."""
    tr = np.zeros((2 * k, n), dtype=np.uint64)
    res = []
    for j in range(k):
        a = b = j + 1
        for i in range(n):
            tr[2 * j, i], tr[2 * j + 1, i] = a, b
            a = (a + b) % P
            b = (b + a) % P
        res.append(int(tr[2 * j + 1, n - 1]))
    A = AirBuilder(2 * k)
    A.pub = res
    for j in range(k):
        A.constraint(A.sub(A.nxt(2 * j), A.add(A.cur(2 * j), A.cur(2 * j + 1))), 1)
        A.constraint(A.sub(A.nxt(2 * j + 1), A.add(A.cur(2 * j + 1), A.nxt(2 * j))), 1)
        A.assert_single(2 * j, 0, j + 1)
        A.assert_single(2 * j + 1, 0, j + 1)
        A.assert_single(2 * j + 1, n - 1, res[j])
    return A.build(), tr


def mulfib2(n):
    """examples/src/fibonacci/mulfib2/air.rs over f64: s0' = s0*s1, s1' = s1*s0' (degree 2), start (1, 2)."""
    tr = np.zeros((2, n), dtype=np.uint64)
    a, b = 1, 2
    for i in range(n):
        tr[0, i], tr[1, i] = a, b
        a = a * b % P
        b = b * a % P
    A = AirBuilder(2)
    A.pub = [int(tr[0, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.mul(A.cur(0), A.cur(1))), 2)
    A.constraint(A.sub(A.nxt(1), A.mul(A.cur(1), A.nxt(0))), 2)
    A.assert_single(0, 0, 1)
    A.assert_single(1, 0, 2)
    A.assert_single(0, n - 1, int(tr[0, n - 1]))
    return A.build(), tr


def periodic_mix(n, cycle=8):
    """Periodic columns + a periodic assertion + degree 3 + two transition exemptions:
         s0' = s0 * k0 + k1          (k0, k1 periodic with cycle `cycle` and 4: degree 1 + 1 cycle)
         s1' = s1^3 + s0 + k1        (degree 3, one cycle)
       column 2 is a flag column equal to 1 on every `cycle`-th step (periodic assertion) with the
       constraint flag * (flag - 1) = 0 (degree 2)."""
    k0 = [(3 * i + 1) % P for i in range(cycle)]
    k1 = [5, 7, 11, 13]
    tr = np.zeros((3, n), dtype=np.uint64)
    a, b = 3, 4
    for i in range(n):
        tr[0, i], tr[1, i], tr[2, i] = a, b, 1 if i % cycle == 0 else 0
        a2 = (a * k0[i % cycle] + k1[i % 4]) % P
        b = (pow(b, 3, P) + a + k1[i % 4]) % P
        a = a2
    A = AirBuilder(3)
    A.periodic = [k0, k1]
    A.exemptions = 2
    A.pub = [int(tr[0, n - 1]), int(tr[1, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.add(A.mul(A.cur(0), A.per(0)), A.per(1))), 1, [cycle])
    b3 = A.mul(A.mul(A.cur(1), A.cur(1)), A.cur(1))
    A.constraint(A.sub(A.nxt(1), A.add(A.add(b3, A.cur(0)), A.per(1))), 3, [4])
    one = A.const(1)
    A.constraint(A.mul(A.cur(2), A.sub(A.cur(2), one)), 2)
    A.assert_single(0, 0, 3)
    A.assert_single(1, 0, 4)
    A.assert_periodic(2, 0, cycle, 1)
    # with two exemptions the last TWO transitions are not enforced; the final values are still asserted
    A.assert_single(0, n - 1, int(tr[0, n - 1]))
    A.assert_single(1, n - 1, int(tr[1, n - 1]))
    return A.build(), tr


PERM_RAP_AUX_WIDTH = 3


def perm_rap(n, seed=5, dyn_last_q=False):
    """Two-segment AIR in the style of examples/src/rescue_raps (a randomised AIR with preprocessing):
    main columns x0, x1 (the FibSmall pair) and b = a permutation of x0's first n-1 values; the aux
    segment proves the permutation with a running product and carries a running sum that mixes main
    columns, a periodic column and both random elements:
        p' * (b + gamma) = p * (x0 + gamma),   p[0] = p[n-1] = 1
        q' = q + alpha * k * x1 * p,           q[0] = 0
        c' = c + 1,                            c[1 + i n/4] asserted as a SEQUENCE (values in E)
    Returns (description, main trace, aux builder(rand [2, d]) -> [2, n, d]).
    dyn_last_q: also assert q[n-1], whose value depends on the random elements (Air::get_aux_assertions(aux_rand_elements),
    air/src/air/mod.rs:279): the description carries a placeholder and `builder.values_fn(rand, values)` fills it in
    (values: [builder.num_values, d], description order); for wf_prove_air_aux_dyn / the oracle's *_dyn entry points."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    tr = np.zeros((3, n), dtype=np.uint64)
    a = b = 1
    for i in range(n):
        tr[0, i], tr[1, i] = a, b
        a = (a + b) % P
        b = (b + a) % P
    perm = rng.permutation(n - 1)
    tr[2, : n - 1] = tr[0, perm]
    tr[2, n - 1] = 12345
    k = [1, 2, 3, 4]
    A = AirBuilder(3)
    A.periodic = [k]
    A.pub = [int(tr[1, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.add(A.cur(0), A.cur(1))), 1)
    A.constraint(A.sub(A.nxt(1), A.add(A.cur(1), A.nxt(0))), 1)
    A.assert_single(0, 0, 1)
    A.assert_single(1, 0, 1)
    A.assert_single(1, n - 1, int(tr[1, n - 1]))
    X = A.aux(PERM_RAP_AUX_WIDTH, 2)
    gamma, alpha = X.rnd(0), X.rnd(1)
    lhs = X.mul(X.anxt(0), X.add(X.cur(2), gamma))
    rhs = X.mul(X.acur(0), X.add(X.cur(0), gamma))
    X.constraint(X.sub(lhs, rhs), 2)
    term = X.mul(X.mul(alpha, X.per(0)), X.mul(X.cur(1), X.acur(0)))
    X.constraint(X.sub(X.anxt(1), X.add(X.acur(1), term)), 2, [4])
    X.assert_single(0, 0, (1, 0, 0))
    X.assert_single(0, n - 1, (1, 0, 0))
    X.assert_single(1, 0, (0, 0, 0))
    if dyn_last_q:
        X.assert_single(1, n - 1, (0, 0, 0))   # placeholder: the running sum's last value is a function of gamma and alpha
    one = X.const(1)
    X.constraint(X.sub(X.anxt(2), X.add(X.acur(2), one)), 1)
    X.assert_sequence(2, 1, n // 4, [(5 + 1 + k * (n // 4), 0, 0) for k in range(4)])

    def builder_py(rand):  # reference implementation of the aux columns, element by element
        d = rand.shape[1]
        g, al = rand[0], rand[1]
        emb = lambda v: np.array([int(v)] + [0] * (d - 1), dtype=np.uint64)
        eadd = lambda x, y: np.array([(int(x[i]) + int(y[i])) % P for i in range(d)], dtype=np.uint64)
        aux = np.zeros((PERM_RAP_AUX_WIDTH, n, d), dtype=np.uint64)
        p, q = emb(1), emb(0)
        for i in range(n):
            aux[0, i], aux[1, i], aux[2, i] = p, q, emb(5 + i)
            q = eadd(q, O.ext_mul(O.ext_mul(al, emb(k[i % 4] * int(tr[1, i]) % P)), p))
            p = O.ext_mul(O.ext_mul(p, eadd(emb(tr[0, i]), g)), O.ext_inv(eadd(emb(tr[2, i]), g)))
        return aux

    def builder(rand):  # the same columns from the C helper (oracle/wf_prover.cpp wfo_perm_rap_aux)
        return O.perm_rap_aux(tr, rand)

    def values_fn(rand, values):  # get_aux_assertions: values in description order [p0, p_last, q0, q_last, 4 sequence values]
        out = values.copy()
        out[3] = builder(rand)[1, n - 1]
        return out

    builder.reference = builder_py
    builder.values_fn = values_fn
    builder.num_values = 8 if dyn_last_q else 7
    return A.build(), tr, builder


def sequence_mix(n, stride=4):
    """Sequence assertions (Assertion::sequence): every `stride`-th value of x0 starting at step 1 is asserted
    (n / stride values: a LargePoly constraint for n / stride >= 63, a SmallPoly one below,
    prover/src/constraints/evaluator/boundary.rs:340,389), next to single and periodic assertions that
    share and do not share its divisor."""
    tr = np.zeros((3, n), dtype=np.uint64)
    a, b = 2, 3
    for i in range(n):
        tr[0, i], tr[1, i], tr[2, i] = a, b, 7 if i % stride == 1 else (i % 5)
        a = (a * a + b) % P
        b = (b + a) % P
    A = AirBuilder(3)
    A.pub = [int(tr[1, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.add(A.mul(A.cur(0), A.cur(0)), A.cur(1))), 2)
    A.constraint(A.sub(A.nxt(1), A.add(A.cur(1), A.nxt(0))), 1)
    A.assert_single(0, 0, 2)
    A.assert_single(1, 0, 3)
    A.assert_single(1, n - 1, int(tr[1, n - 1]))
    A.assert_sequence(0, 1, stride, [int(v) for v in tr[0, 1::stride]])   # first_step != 0: x offset g^-1
    A.assert_periodic(2, 1, stride, 7)                                    # same divisor as the sequence
    A.assert_sequence(2, 0, n // 2, [int(tr[2, 0]), int(tr[2, n // 2])])  # two values, first_step 0 (no cell asserted twice)
    return A.build(), tr


def rescue_like(n, width=6):
    """A heavy single-segment AIR for the generic evaluator (the shape of a Rescue-Prime half round,
    examples/src/rescue/rescue.rs: S-box x^7, MDS layer, round constants): state' = M * state^7 + k with a circulant
    M of small entries; `width` constraints of degree 7 (constraint-evaluation blowup 8), ~19 field operations per column
    and row, 2 * width + ~19 * width registers in the transition program."""
    m_row = [2, 3, 1, 1, 5, 7, 1, 4][:width]
    ark = [(0x9E3779B97F4A7C15 * (j + 1)) % P for j in range(width)]
    tr = np.zeros((width, n), dtype=np.uint64)
    st = [j + 1 for j in range(width)]
    for i in range(n):
        for j in range(width):
            tr[j, i] = st[j]
        p7 = [pow(v, 7, P) for v in st]
        st = [(sum(m_row[(c - j) % width] * p7[c] for c in range(width)) + ark[j]) % P for j in range(width)]
    A = AirBuilder(width)
    A.pub = [int(tr[0, n - 1])]
    mc = [A.const(v) for v in m_row]
    kc = [A.const(v) for v in ark]
    p7 = []
    for c in range(width):
        x = A.cur(c)
        x2 = A.mul(x, x)
        x4 = A.mul(x2, x2)
        x3 = A.mul(x2, x)
        p7.append(A.mul(x3, x4))
    for j in range(width):
        acc = kc[j]
        for c in range(width):
            acc = A.add(acc, A.mul(mc[(c - j) % width], p7[c]))
        A.constraint(A.sub(A.nxt(j), acc), 7)
    for j in range(width):
        A.assert_single(j, 0, j + 1)
    A.assert_single(0, n - 1, int(tr[0, n - 1]))
    return A.build(), tr
