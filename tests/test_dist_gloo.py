"""world_size-2 and -4 `gloo` tests of the multi-GPU host logic (winterfell_b200/dist.py) on CPU: the sharding
plan, the all-to-all row re-sharding, the subtree-root all-gather and the top-of-tree merge. Local
compute is done by a CPU test backend built on the oracle; the result must equal the single-device
commitment of the whole trace."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """CPU stand-in for CudaBackend (row-major exchange format)."""

    def pack(self, cols, ncols, n, log_blowup, world):
        from oracle import oracle as o
        polys = o.interpolate_columns(cols.numpy().view(np.uint64))
        rows = torch.from_numpy(o.lde_rows(polys, 1 << log_blowup).view(np.int64))
        return rows.view(world, -1)

    def commit(self, hash_id, recv, rows_per, ncols_total, cl):
        from oracle import oracle as o
        world = recv.shape[0]
        rows = recv.view(world, rows_per, cl).permute(1, 0, 2).reshape(rows_per, ncols_total).contiguous()
        lv = o.hash_rows(hash_id, rows.numpy().view(np.uint64))
        nd = o.merkle_nodes(hash_id, lv)
        return nd[1].tobytes(), rows, torch.from_numpy(lv), torch.from_numpy(nd)


def _worker(rank, world, port, hash_id, log_n, cols, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from winterfell_b200 import dist as wd
    n = 1 << log_n
    trace = o.rand_elems((cols, n), 99)            # same seed on every rank: the whole trace
    lo, hi = wd.column_range(cols, world, rank)
    local = torch.from_numpy(trace[lo:hi].copy().view(np.int64))
    root, rows, digests, nodes = wd.sharded_trace_commit(OracleBackend(), hash_id, local, cols, log_n, 3)
    # single-device reference: commitment of the full trace
    lde = o.lde_rows(o.interpolate_columns(trace), 8)
    want_nodes = o.merkle_nodes(hash_id, o.hash_rows(hash_id, lde))
    N = n * 8
    ok = root == want_nodes[1].tobytes()
    ok = ok and (rows.numpy().view(np.uint64) == lde[rank * N // world:(rank + 1) * N // world]).all()
    # my subtree is the heap node world + rank of the full tree
    ok = ok and nodes.numpy()[1].tobytes() == want_nodes[world + rank].tobytes()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("hash_id,log_n,cols,world", [(0, 6, 4, 2), (1, 5, 2, 2), (0, 5, 8, 4)])
def test_sharded_trace_commit_gloo(hash_id, log_n, cols, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, hash_id, log_n, cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


def test_column_ranges_and_top_levels(oracle):
    from winterfell_b200 import dist as wd
    assert [wd.column_range(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert [wd.column_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    leaves = np.random.default_rng(1).integers(0, 256, (8, 32), dtype=np.uint8)
    nodes = oracle.merkle_nodes(0, leaves)
    # the four depth-2 nodes (heap 4..7) merge to the root
    assert wd.top_levels(0, [nodes[4 + i].tobytes() for i in range(4)]) == nodes[1].tobytes()


# ---- sharded openings (wf_prove_fib_sharded): every rank holds one subtree; a batch opening is gathered from the owners,
#      summed over the ranks (gloo all_reduce), the top levels patched in from the all-gathered roots ----
def _open_worker(rank, world, port, hash_id, log_leaves, positions, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import winterfell_b200 as wf
    from oracle import oracle as o
    from winterfell_b200 import dist as wd
    n = 1 << log_leaves
    leaves = np.random.default_rng(7).integers(0, 256, (n, 32), dtype=np.uint8)   # same on every rank
    if hash_id == 1:  # Rp64 digests are four canonical field elements
        leaves = (leaves.view(np.uint64) % np.uint64(0xFFFFFFFF00000001)).view(np.uint8).reshape(n, 32)
    n_local = n // world
    mine = leaves[rank * n_local:(rank + 1) * n_local]
    sub = o.merkle_nodes(hash_id, mine)                                             # my subtree (heap, root at 1)
    # all-gather of the subtree roots, top levels on every rank (ShardTree in prover.cu)
    root_t = torch.from_numpy(sub[1].copy())
    roots = [torch.empty_like(root_t) for _ in range(world)]
    dist.all_gather(roots, root_t)
    top = [None] * (2 * world)
    for i in range(world):
        top[world + i] = roots[i].numpy().tobytes()
    for i in range(world - 1, 0, -1):
        top[i] = wf.host_merge(hash_id, top[2 * i], top[2 * i + 1])
    want, idx = wf.sharded_opening_plan(n, world, rank, np.array(positions, dtype=np.uint64))
    got = np.zeros((len(want), 32), dtype=np.uint8)
    NONE, TOP = np.uint64(2**64 - 1), np.uint64(2**64 - 2)
    for s, (e, li) in enumerate(zip(want, idx)):
        if li == NONE or li == TOP:
            continue
        got[s] = sub[int(li)] if li < n_local else mine[int(li) - n_local]
    t = torch.from_numpy(got.view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                                         # disjoint owners: the sum is the gather
    got = t.numpy().view(np.uint8).reshape(len(want), 32).copy()
    for s, li in enumerate(idx):
        if li == TOP:
            got[s] = np.frombuffer(top[int(want[s])], dtype=np.uint8)
    # single-device reference: the same digests straight from the full tree
    full = o.merkle_nodes(hash_id, leaves)
    ref = np.stack([full[int(e)] if e < n else leaves[int(e) - n] for e in want])
    ok = (got == ref).all() and top[1] == full[1].tobytes()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("hash_id,log_leaves,world,positions", [(0, 8, 2, [0, 1, 77, 128, 255]), (0, 10, 4, [5, 300, 301, 767, 1023, 512]),
                                                               (1, 6, 2, [3, 40]), (0, 6, 8, [0, 9, 63, 31, 32])])
def test_sharded_opening_gloo(hash_id, log_leaves, world, positions):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_open_worker, args=(r, world, port, hash_id, log_leaves, positions, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
