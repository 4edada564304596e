"""wf_air_check: the host-side checks of an AIR description (structure, degrees vs blowup, periodic columns, assertion
validity / overlaps: the conditions the reference panics on in Air::new, BoundaryConstraints::new, prepare_assertions —
air/src/air/boundary/mod.rs:190-215, air/src/air/assertions/mod.rs:62-230). Runs without a GPU; the fuzz part feeds the
parser truncated, extended and randomly mutated descriptions: it must answer WF_OK or WF_ERR_INVALID, never crash or hang."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import airs  # noqa: E402
import winterfell_b200 as wf  # noqa: E402

WF_OK, WF_ERR_INVALID = 0, -2


def _all():
    out = [("mulfib2", airs.mulfib2(64)[0], 6), ("periodic_mix", airs.periodic_mix(64)[0], 6), ("sequence_mix", airs.sequence_mix(64)[0], 6),
           ("rescue_like", airs.rescue_like(64)[0], 6), ("fib_small_x", airs.fib_small_x(4, 128)[0], 7), ("perm_rap", airs.perm_rap(128)[0], 7)]
    return out


def test_valid_descriptions_pass_and_reasons_are_named():
    for name, d, log_n in _all():
        rc, msg = wf.air_check(d, log_n, 8)
        assert (rc, msg) == (WF_OK, ""), (name, rc, msg)
    d = airs.periodic_mix(64)[0]
    assert wf.air_check(d, 6, 2) == (WF_ERR_INVALID, "blowup factor too small for the constraint degrees")   # degree 3 needs blowup >= 4
    assert wf.air_check(d, 2, 8)[0] == WF_ERR_INVALID                                                          # trace shorter than 8 rows
    assert wf.air_check(d, 6, 12)[0] == WF_ERR_INVALID                                                         # blowup not a power of two
    d = airs.sequence_mix(64)[0]
    assert wf.air_check(d, 5, 8) == (WF_ERR_INVALID, "invalid assertion")        # n / stride no longer equals the number of values
    d = airs.periodic_mix(64)[0]
    assert wf.air_check(d, 3, 8)[0] == WF_ERR_INVALID                                                          # 8-row trace: cycles / columns too long


def test_overlapping_assertions_are_named():
    # two assertions on the same cell (Assertion::overlaps_with, assertions/mod.rs:175-208; prepare_assertions panics,
    # boundary/mod.rs:205-210): the same two descriptions the GPU test proves with
    n = 64
    _, trace = airs.mulfib2(n)
    A = airs.AirBuilder(2)
    A.pub = [int(trace[0, n - 1])]
    A.constraint(A.sub(A.nxt(0), A.mul(A.cur(0), A.cur(1))), 2)
    A.constraint(A.sub(A.nxt(1), A.mul(A.cur(1), A.nxt(0))), 2)
    A.assert_single(0, 0, 1)
    A.assert_single(1, 0, 2)
    A.assert_sequence(1, 0, n // 2, [2, int(trace[1, n // 2])])   # step 0 of column 1 is asserted twice
    rc, msg = wf.air_check(A.build(), 6, 8)
    assert rc == WF_ERR_INVALID and "overlaps" in msg
    B = airs.AirBuilder(2)
    B.pub = A.pub
    B.constraint(B.sub(B.nxt(0), B.mul(B.cur(0), B.cur(1))), 2)
    B.constraint(B.sub(B.nxt(1), B.mul(B.cur(1), B.nxt(0))), 2)
    B.assert_single(0, 0, 1)
    B.assert_periodic(1, 0, 4, 2)
    B.assert_single(1, 8, 5)                                         # step 8 = 0 + 2 * 4 is covered by the periodic one
    rc, msg = wf.air_check(B.build(), 6, 8)
    assert rc == WF_ERR_INVALID and "overlaps" in msg
    B2 = airs.AirBuilder(2)
    B2.pub = A.pub
    B2.constraint(B2.sub(B2.nxt(0), B2.mul(B2.cur(0), B2.cur(1))), 2)
    B2.constraint(B2.sub(B2.nxt(1), B2.mul(B2.cur(1), B2.nxt(0))), 2)
    B2.assert_single(0, 0, 1)
    B2.assert_periodic(1, 0, 4, 2)
    B2.assert_single(1, 9, 5)                                        # step 9 is not on the periodic assertion's grid
    assert wf.air_check(B2.build(), 6, 8) == (WF_OK, "")


@pytest.mark.parametrize("seed", range(6))
def test_fuzzed_descriptions_never_crash(seed):
    rng = np.random.default_rng(1000 + seed)
    interesting = np.array([0, 1, 2, 3, 4, 7, 8, 16, 255, 256, 4096, 1 << 20, (1 << 32) - 1, 1 << 32, (1 << 63), wf.P - 1, wf.P, (1 << 64) - 1], dtype=np.uint64)
    seen = {WF_OK: 0, WF_ERR_INVALID: 0}
    for name, d, log_n in _all():
        for _ in range(400):
            m = d.copy()
            kind = rng.integers(0, 5)
            if kind == 0:                                   # truncate
                m = m[: rng.integers(0, len(m))]
            elif kind == 1:                                 # extend with junk
                m = np.concatenate([m, rng.choice(interesting, size=rng.integers(1, 9))])
            elif kind == 2:                                 # overwrite a few words with boundary values
                for i in rng.integers(0, len(m), size=rng.integers(1, 4)):
                    m[i] = rng.choice(interesting)
            elif kind == 3:                                 # small perturbations (off-by-one counts, indices)
                for i in rng.integers(0, len(m), size=rng.integers(1, 4)):
                    m[i] = np.uint64((int(m[i]) + int(rng.integers(-2, 3))) % (1 << 64))
            else:                                           # random words
                for i in rng.integers(0, len(m), size=rng.integers(1, 6)):
                    m[i] = np.uint64(int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2)))
            rc, msg = wf.air_check(np.ascontiguousarray(m, dtype=np.uint64), int(rng.integers(3, 12)), int(rng.choice([2, 4, 8, 16, 128])))
            assert rc in (WF_OK, WF_ERR_INVALID), (name, rc, msg)
            assert (rc == WF_OK) == (msg == "")
            seen[rc] += 1
    assert seen[WF_ERR_INVALID] > 500 and seen[WF_OK] > 0      # both outcomes were exercised
