"""CPU (gloo, world size 2 and 4): the multi-GPU host logic -- column-sharded LDE, all-to-all to row blocks, per-rank
subtrees, cap gather and top-level compression -- yields the single-device root bit-for-bit.  The arithmetic backend is
the oracle here (the production backend is GpuOps); only the sharding/collective logic is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from zkvm_prover_b200 import dist as D


class OracleOps:
    def to_device(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint32).view(np.int32))

    def lde(self, local, added_bits, shift):
        a = local.numpy().view(np.uint32)
        return torch.from_numpy(O.coset_lde_batch(a, added_bits, shift, bitrev_out=True).view(np.int32))

    def subtree_root(self, chunks):
        root, _ = O.merkle_commit([c.contiguous().numpy().view(np.uint32) for c in chunks])
        return root

    def compress(self, pairs):
        return O.compress_pairs(pairs)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, w, added_bits, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = O.fill(n * w, 1234).reshape(n, w)
        wg = w // world
        local = np.ascontiguousarray(full[:, rank * wg:(rank + 1) * wg])
        shift = int(O.to_monty([31])[0])
        root, cap = D.sharded_lde_commit(OracleOps(), local, added_bits, shift)
        np.save(os.path.join(out_dir, f"root{rank}.npy"), root)
        np.save(os.path.join(out_dir, f"cap{rank}.npy"), cap)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,w", [(2, 64, 16), (4, 32, 8), (2, 32, 64)])   # the last one runs 4 column strips per rank
def test_sharded_lde_commit_matches_single_device(tmp_path, world, n, w):
    added_bits = 1
    mp.spawn(_worker, args=(world, _free_port(), n, w, added_bits, str(tmp_path)), nprocs=world, join=True)
    full = O.fill(n * w, 1234).reshape(n, w)
    lde = O.coset_lde_batch(full, added_bits, int(O.to_monty([31])[0]), bitrev_out=True)
    root, layers = O.merkle_commit([lde])
    lg = world.bit_length() - 1
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"root{r}.npy"), root)
        assert np.array_equal(np.load(tmp_path / f"cap{r}.npy"), layers[len(layers) - 1 - lg])


def test_segment_assignment_and_cap():
    assert D.segment_assignment(5, 2) == [[0, 2, 4], [1, 3]]
    assert D.segment_assignment(3, 4) == [[0], [1], [2], []]
    cap = O.fill(64, 9).reshape(8, 8)
    lay = cap
    while lay.shape[0] > 1:
        lay = O.compress_pairs(lay.reshape(-1, 16))
    assert np.array_equal(D.combine_cap(cap, O.compress_pairs), lay[0])
