"""Multi-GPU plumbing for the column-sharded trace commitment (option 1).

One process per GPU (`torch.distributed`, backend "nccl"; the host-logic tests run the same code over
"gloo" on CPU with a test backend). The reference has no distributed code at all — its only
distribution-aware feature is `PartitionOptions` (air/src/options.rs:405-445) — so the decomposition
is new, but the result is bit-identical to the single-device commitment (`DefaultTraceLde::new`,
prover/src/trace/trace_lde/default/mod.rs:63, :245-282):

  1. rank g owns columns [g*c/G, (g+1)*c/G): interpolate + LDE locally (no communication; K1/K2 are
     independent per column).
  2. ONE all-to-all turns column shards into row shards: rank r receives rows [r*N/G, (r+1)*N/G) of
     every column block (each rank sends (G-1)/G of its LDE slice).
  3. rank r hashes its rows and builds the Merkle subtree over them — rows [r*N/G, (r+1)*N/G) are
     exactly one depth-log2(G) subtree of the reference's heap layout (crypto/src/merkle/mod.rs:344-368).
  4. ONE all-gather of the G subtree roots (32 bytes each); every rank computes the top log2(G) levels
     redundantly on the host (H::merge, crypto/src/hash/mod.rs:45).

This module contains no arithmetic: local compute goes through a backend (the CUDA context in the
product; the tests substitute a CPU backend), collectives through torch.distributed.
"""
import numpy as np
import torch
import torch.distributed as dist

import winterfell_b200 as wf


def column_range(ncols, world, rank):
    """Contiguous column block of `rank`; the first (ncols % world) ranks get one extra column."""
    base, extra = divmod(ncols, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def top_levels(hash_id, roots):
    """Root of the tree whose depth-log2(G) nodes are `roots` (list of 32-byte digests, in order).
    Same heap rule as build_merkle_nodes: parent(i) = merge(child 2i, child 2i+1)."""
    level = list(roots)
    assert len(level) & (len(level) - 1) == 0 and len(level) >= 1
    while len(level) > 1:
        level = [wf.host_merge(hash_id, level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    return level[0]


class CudaBackend:
    """Local compute on this rank's GPU through the C ABI.

    Exchange format (when the rank owns a multiple of 8 columns, i.e. whole segments): the LDE is written
    by the library straight into a torch buffer in segment layout [local segment][N rows][8]; the send
    buffer is that tensor regrouped by destination, [dest][local segment][rows_per][8] (one device copy);
    what arrives, [source][local segment][rows_per][8], IS the segment layout of the rank's
    rows_per x ncols_total row shard (global segment = source * local_segments + local segment), so the
    commitment kernels run on the receive buffer in place. Other widths take a row-major exchange."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", torch.cuda.current_device())

    def pack(self, cols_dev, ncols, n, log_blowup, world):
        """cols_dev: int64 CUDA tensor [ncols, n]. Returns the all-to-all send tensor [world, chunk]."""
        N = n << log_blowup
        m = self.ctx.mat_from_device_columns(cols_dev.data_ptr(), ncols, n)
        polys = m.interpolate()
        if ncols % 8 == 0:
            nsl = ncols // 8
            seg = torch.empty((nsl, N, 8), dtype=torch.int64, device=self.device)
            out = self.ctx.mat_wrap_device(seg.data_ptr(), N, ncols)
            polys.lde_into(log_blowup, out)
            self.ctx.sync()
            for h in (m, polys, out):
                h.free()
            rp = N // world
            return seg.view(nsl, world, rp * 8).permute(1, 0, 2).contiguous().view(world, nsl * rp * 8)
        lde = polys.lde(log_blowup)
        rows = torch.empty((N, ncols), dtype=torch.int64, device=self.device)
        lde.to_device_rows(rows.data_ptr())
        self.ctx.sync()
        for h in (m, polys, lde):
            h.free()
        return rows.view(world, (N // world) * ncols)

    def commit(self, hash_id, recv, rows_per, ncols_total, cl):
        """recv: [world, chunk] as received. Returns (subtree root bytes, row shard tensor, leaf digests, nodes)."""
        world = recv.shape[0]
        if cl % 8 == 0:
            shard = self.ctx.mat_wrap_device(recv.data_ptr(), rows_per, ncols_total)
            tree = self.ctx.commit_rows(hash_id, shard)
            root = tree.root()
            leaves, nodes = tree.to_host() if rows_per <= (1 << 16) else (None, None)
            shard.free()
            tree.free()
            return root, recv, leaves, nodes
        rows = recv.view(world, rows_per, cl).permute(1, 0, 2).reshape(rows_per, ncols_total).contiguous()
        digests = torch.empty(rows_per * 32, dtype=torch.uint8, device=self.device)
        nodes = torch.empty(rows_per * 32, dtype=torch.uint8, device=self.device)
        self.ctx.hash_rows_dev(hash_id, rows.data_ptr(), rows_per, ncols_total, digests.data_ptr())
        self.ctx.merkle_dev(hash_id, digests.data_ptr(), rows_per, nodes.data_ptr())
        self.ctx.sync()
        return bytes(nodes[32:64].cpu().numpy()), rows, digests, nodes


def sharded_trace_commit(backend, hash_id, local_cols, ncols_total, log_n, log_blowup, group=None):
    """Column-sharded trace commitment. `local_cols`: this rank's column block [c_local, n] as a
    torch int64 tensor on the backend's device. All ranks must own the same number of columns
    (ncols_total % world == 0). Returns (root, row shard, subtree leaf digests, subtree nodes)."""
    world = dist.get_world_size(group)
    n = 1 << log_n
    N = n << log_blowup
    assert ncols_total % world == 0, "columns must divide evenly across ranks"
    assert N % world == 0 and world & (world - 1) == 0, "world size must be a power of two"
    cl = ncols_total // world
    assert tuple(local_cols.shape) == (cl, n)
    rows_per = N // world
    # 1. local LDE of the owned columns, regrouped by destination rank
    send = backend.pack(local_cols, cl, n, log_blowup, world)
    # 2. ONE all-to-all: row range r of every column block goes to rank r
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    # 3. leaves + subtree over my rows
    root_local, rows, digests, nodes = backend.commit(hash_id, recv, rows_per, ncols_total, cl)
    # 4. all-gather the subtree roots, finish the top of the tree on every rank
    mine = torch.frombuffer(bytearray(root_local), dtype=torch.uint8).to(send.device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    roots = [bytes(g.cpu().numpy()) for g in gathered]
    return top_levels(hash_id, roots), rows, digests, nodes
