"""Multi-GPU layer: only where the path shards.

* segments / independent proofs: `segment_assignment` -- replicas, no data-path collective; roots gathered at the end.
* one wide matrix: `sharded_lde_commit` -- columns sharded for the LDE (columns are independent), one all-to-all to
  row blocks, per-rank subtree over its contiguous (bit-reversed-order) rows, all-gather of the G subtree roots (the
  "cap"), the top log2(G) levels compressed redundantly on every rank.  The root is bit-identical to the single-GPU root.

The collective plumbing is torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).  The arithmetic comes
from an `ops` object: `GpuOps` (the CUDA library) in production; the CPU tests inject an oracle-backed one.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def segment_assignment(n_segments: int, world: int):
    """round-robin continuation segments (or chunk tasks) over ranks: independent units, no collective"""
    return [[s for s in range(n_segments) if s % world == r] for r in range(world)]


def _all_to_all_equal(recv: torch.Tensor, send: torch.Tensor, group=None):
    """recv[s] <- block `rank` of rank s's `send` (equal splits along dim 0)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    try:
        if dist.get_backend(group) != "gloo":
            dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
            return
    except Exception:
        pass
    ops = []
    recv[rank].copy_(send[rank])
    for peer in range(world):
        if peer == rank:
            continue
        ops.append(dist.P2POp(dist.isend, send[peer].contiguous(), peer, group))
        ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class GpuOps:
    """arithmetic backend = libb200zk on this rank's GPU; tensors are int32 CUDA tensors viewed as BabyBear u32"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def to_device(self, arr) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint32).view(np.int32)).to(self.device)

    def lde(self, local: torch.Tensor, added_bits: int, shift: int) -> torch.Tensor:
        n, w = local.shape
        out = torch.empty((n << added_bits, w), dtype=torch.int32, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        src = self.ctx.wrap(local.data_ptr(), n, w, keepalive=local)
        dst = self.ctx.wrap(out.data_ptr(), n << added_bits, w, keepalive=out)
        self.ctx.check(self.ctx.lib.b200zk_coset_lde_batch_into(self.ctx.h, src.h, added_bits, shift, 1, dst.h))
        self.ctx.sync()
        return out

    def subtree_root(self, chunks) -> np.ndarray:
        torch.cuda.current_stream(self.device).synchronize()
        mats = [self.ctx.wrap(c.data_ptr(), c.shape[0], c.shape[1], keepalive=c) for c in chunks]
        arr = (C.c_void_p * len(mats))(*[m.h for m in mats])
        root = np.empty(8, np.uint32)
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_merkle_commit(self.ctx.h, arr, len(mats), 0, root.ctypes.data, C.byref(t)))
        self.ctx.lib.b200zk_tree_free(self.ctx.h, t)
        return root

    def compress(self, pairs: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 16)
        out = np.empty((p.shape[0], 8), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_compress_pairs(self.ctx.h, p.ctypes.data, out.ctypes.data, p.shape[0]))
        return out


def combine_cap(cap: np.ndarray, compress) -> np.ndarray:
    """cap: (G, 8) subtree roots in rank order (G a power of two) -> Merkle root of the whole tree"""
    layer = np.ascontiguousarray(cap, dtype=np.uint32)
    while layer.shape[0] > 1:
        layer = compress(layer.reshape(-1, 16))
    return layer[0]


def _all_to_all_start(recv: torch.Tensor, send: torch.Tensor, group=None):
    """begin recv[s] <- block `rank` of rank s's `send`; returns a waitable (None when it already completed)"""
    try:
        if dist.get_backend(group) != "gloo":
            return dist.all_to_all_single(recv.view(-1), send.view(-1), group=group, async_op=True)
    except Exception:
        pass
    _all_to_all_equal(recv, send, group)
    return None


def sharded_lde_commit(ops, local_cols, added_bits: int, shift: int, group=None, strips: int = 4):
    """Every rank holds a column shard (N x W/G, same N and W/G everywhere) of one trace matrix.
    Returns (root[8] -- identical on all ranks and to the single-GPU commit of the full LDE --, cap (G,8)).

    The shard is extended in `strips` column strips; the all-to-all of strip j (NCCL, its own stream) runs while strip j+1 is
    being extended, so only the last strip's exchange is exposed before the hashing starts.  The sponge absorbs the
    columns in global order: rank 0's strips first, then rank 1's, ..."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world & (world - 1):
        raise ValueError("the number of ranks must be a power of two")
    local = local_cols if isinstance(local_cols, torch.Tensor) else ops.to_device(local_cols)
    wg = local.shape[1]
    if strips < 1 or wg % strips or (wg // strips) % 8:
        strips = 1                                          # keep every exchanged block a whole number of sponge chunks
    ws = wg // strips
    recvs, pending = [], []
    for j in range(strips):
        part = local if strips == 1 else local[:, j * ws:(j + 1) * ws].contiguous()
        lde = ops.lde(part, added_bits, shift)              # (M, ws), rows in bit-reversed order; returns when it is complete
        m = lde.shape[0]
        if m % world:
            raise ValueError("LDE height must be divisible by the number of ranks")
        mg = m // world
        send = lde.view(world, mg, ws)                      # row block r goes to rank r
        recv = torch.empty_like(send)
        pending.append((_all_to_all_start(recv, send, group), send))   # keep `send` alive until the exchange is done
        recvs.append(recv)                                  # recv[s] = my row block of rank s's strip j
    for work, _ in pending:
        if work is not None:
            work.wait()
    chunks = [recvs[j][s] for s in range(world) for j in range(strips)]   # global column order
    root_local = ops.subtree_root(chunks)                   # one sponge over all chunks, then the subtree of my rows
    dev = recvs[0].device
    cap_t = [torch.zeros(8, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cap_t, torch.from_numpy(root_local.astype(np.int64)).to(dev), group=group)
    cap = np.stack([c.cpu().numpy().astype(np.uint32) for c in cap_t])
    return combine_cap(cap, ops.compress), cap


class PeerExchange:
    """Receive buffers for `sharded_lde_commit_p2p`: every rank owns one [world][M / world][wg] buffer (cudaMalloc, exported through a
    CUDA IPC handle) and maps the buffers of all other ranks, so the last NTT pass can store finished tiles straight into the
    row-block owner's memory over NVLink.  Build once per shape, reuse for every commit."""

    def __init__(self, ctx, rows_lde: int, wg: int, group=None):
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if rows_lde % self.world:
            raise ValueError("LDE height must be divisible by the number of ranks")
        self.mg, self.wg = rows_lde // self.world, wg
        self.nbytes = 4 * rows_lde * wg
        own, handle = C.c_void_p(), (C.c_uint8 * 64)()
        ctx.check(ctx.lib.b200zk_peer_alloc(ctx.h, self.nbytes, C.byref(own), handle))
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
                continue
            p = C.c_void_p()
            hb = (C.c_uint8 * 64).from_buffer_copy(h)
            ctx.check(ctx.lib.b200zk_peer_open(ctx.h, hb, C.byref(p)))
            self.ptrs.append(p.value)
        self.ptr_array = (C.c_void_p * self.world)(*self.ptrs)

    def chunk(self, sender: int):
        """my row block of `sender`'s columns: an (mg, wg) matrix inside my receive buffer"""
        return self.ctx.wrap(self.own + 4 * sender * self.mg * self.wg, self.mg, self.wg)

    def matrix(self):
        """the rows layout (b200zk_coset_lde_scatter_rows): my whole row block as ONE (mg, world * wg) matrix"""
        return self.ctx.wrap(self.own, self.mg, self.world * self.wg)

    def close(self):
        if getattr(self, "ptrs", None):
            dist.barrier(self.group)   # nobody may still be storing into a buffer that is about to go away
            for r, p in enumerate(self.ptrs):
                if r != self.rank:
                    self.ctx.lib.b200zk_peer_close(self.ctx.h, C.c_void_p(p))
            dist.barrier(self.group)
            self.ctx.lib.b200zk_peer_free(self.ctx.h, C.c_void_p(self.own))
            self.ptrs = None


def sharded_lde_commit_p2p(ctx, local_cols, added_bits: int, shift: int, exch: PeerExchange | None = None, group=None, rows_layout: bool = True):
    """`sharded_lde_commit` with the exchange fused into the transform: b200zk_coset_lde_scatter stores the finished tiles of the last
    NTT pass directly into the owner's receive buffer (TMA over a peer mapping), so there is no all-to-all and no staging copy.
    local_cols: this rank's column shard as a DeviceMatrix (N x wg).  Returns (root, cap) like `sharded_lde_commit`.
    rows_layout (default): the receive buffer is one row-major (M / world) x W matrix (every sender stores its columns at its own
    column offset), so the owner hashes ONE wide matrix with the fast leaf kernel; False keeps one (M / world) x wg slot per sender
    and commits them as `world` matrices of one height (same root: MMCS concatenates their rows)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n, wg = local_cols.rows, local_cols.width
    own_exch = exch is None
    if own_exch:
        exch = PeerExchange(ctx, n << added_bits, wg, group)
    trace = os.environ.get("B200ZK_DIST_TRACE") == "1"   # experiments: host-clock phase times on rank 0 (adds synchronisation)
    marks = []

    def mark(what):
        if trace:
            torch.cuda.synchronize()
            marks.append((what, time.perf_counter()))
    try:
        mark("start")
        dist.barrier(group)            # every rank is done reading its receive buffer from the previous call
        mark("barrier")
        fn = ctx.lib.b200zk_coset_lde_scatter_rows if rows_layout else ctx.lib.b200zk_coset_lde_scatter
        ctx.check(fn(ctx.h, local_cols.h, added_bits, shift, world, rank, exch.ptr_array))
        ctx.sync()                     # my stores (local and remote) are complete ...
        mark("lde + scatter")
        dist.barrier(group)            # ... and so are everybody else's into my buffer
        mark("barrier")
        chunks = [exch.matrix()] if rows_layout else [exch.chunk(s) for s in range(world)]
        arr = (C.c_void_p * len(chunks))(*[m.h for m in chunks])
        root_local = np.empty(8, np.uint32)
        t = C.c_void_p()
        ctx.check(ctx.lib.b200zk_merkle_commit(ctx.h, arr, len(chunks), 0, root_local.ctypes.data, C.byref(t)))
        ctx.lib.b200zk_tree_free(ctx.h, t)
        mark("subtree commit")
        dev = torch.device("cuda", ctx.device)
        cap_t = [torch.zeros(8, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(cap_t, torch.from_numpy(root_local.astype(np.int64)).to(dev), group=group)
        cap = np.stack([c.cpu().numpy().astype(np.uint32) for c in cap_t])
        mark("cap all-gather")
        root = combine_cap(cap, GpuOps(ctx).compress)
        mark("top levels")
        if trace and rank == 0:
            print("[dist] " + "  ".join(f"{w} {1e3 * (b - a):.2f} ms" for (_, a), (w, b) in zip(marks, marks[1:])), flush=True)
        return root, cap
    finally:
        if own_exch:
            exch.close()
