"""Code-generation guards that need no GPU: the objects build() just compiled are disassembled (cuobjdump) and checked for the
properties DESIGN.md's kernel table relies on — sm_100a code, the NTT pass stages its twiddle tables with a TMA bulk copy and
keeps its tile in registers / shared memory (no local-memory traffic), 128-bit global and shared accesses, the field
multiplication's (2^32 - 1) multiplier comes from the constant bank (IMAD.WIDE, not IMAD.HI pairs), the BLAKE3 row kernels
hold state and message in registers."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "winterfell_b200", "_build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(os.path.join(BUILD, "ntt2.o")) and os.path.exists(CUOBJDUMP)),
                                reason="objects not built (python -c 'import __graft_entry__ as g; g.build()') or no cuobjdump")


def _functions(obj):
    out = subprocess.run([CUOBJDUMP, "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
    fns, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            fns[cur].append(line)
    arch = re.findall(r"arch = (sm_\w+)", out)
    return fns, arch


def _ops(lines):
    ops = []
    for l in lines:
        body = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l)
        body = re.sub(r"^@!?U?P[0-9T]\s+", "", body)
        ops.append(body.split()[0].rstrip(";"))
    return ops


def test_ntt_pass_kernels():
    fns, arch = _functions("ntt2.o")
    assert arch and set(arch) == {"sm_100a"}
    big = {n: l for n, l in fns.items() if "ntt2_pass_kernel" in n and "Li11E" in n}
    assert len(big) == 2, list(fns)                        # strided and contiguous 2^11-point passes (the cfg3 kernels)
    for name, lines in big.items():
        ops = _ops(lines)
        assert any(o.startswith("UBLKCP") for o in ops), name           # TMA bulk copy of the round twiddle table
        assert any(o.startswith("SYNCS") for o in ops), name            # ... completed through an mbarrier
        assert not any(o.startswith(("LDL", "STL")) for o in ops), name  # no spills, no local arrays
        assert any(o.startswith("LDG.E.128") or o.startswith("LDG.E.64") for o in ops), name
        assert any(o.startswith("STS.128") for o in ops) and any(o.startswith("LDS.128") for o in ops), name
        wide = sum(o.startswith("IMAD.WIDE") for o in ops)
        hi = sum(o.startswith("IMAD.HI") for o in ops)
        assert wide > 4 * hi, (name, wide, hi)                           # products as IMAD.WIDE (the half-rate IMAD.HI stays rare)


def test_blake3_row_kernels_stay_in_registers():
    fns, _ = _functions("commit.o")
    for key in ("hash_rows_blake3_w8_kernel", "hash_rows_blake3_w8c8_kernel", "merkle_level_blake3_kernel"):
        hit = {n: l for n, l in fns.items() if key in n}
        assert hit, key
        for name, lines in hit.items():
            ops = _ops(lines)
            assert not any(o.startswith(("LDL", "STL")) for o in ops), name
            # 7 rounds x 8 G functions x 4 rotations, fully unrolled (funnel shifts; byte-aligned rotations may become PRMT)
            assert sum(o.startswith(("SHF", "PRMT")) for o in ops) >= 7 * 8 * 4 - 24, name


def test_rescue_kernels_interleave_six_chains():
    fns, _ = _functions("commit.o")
    hit = [l for n, l in fns.items() if "merkle_level_alg_kernelILi1E" in n]
    assert len(hit) == 1
    lines = hit[0]
    addr = [int(re.match(r"\s+/\*([0-9a-f]+)\*/", l).group(1), 16) for l in lines]
    ops = _ops(lines)
    # the inverse S-box loops carry six independent squarings per iteration: a backward branch whose body (target .. branch)
    # holds the 6 x 3 wide products of six squarings in about a hundred instructions
    loops = []
    for a, o, l in zip(addr, ops, lines):
        m = re.search(r"BRA\S*\s+(?:U?P\d,\s+|!U?P\d,\s+)?0x([0-9a-f]+)", l)
        if o.startswith("BRA") and m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            loops.append([oo for aa, oo in zip(addr, ops) if t <= aa <= a])
    assert any(sum(o.startswith("IMAD.WIDE") for o in b) >= 18 and len(b) < 140 for b in loops), [len(b) for b in loops]
