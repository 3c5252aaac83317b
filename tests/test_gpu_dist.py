"""GPU, >= 2 devices: column-sharded LDE + all-to-all (NCCL) + per-rank subtrees + cap gather gives the single-GPU root."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, w, out_dir):
    import torch
    import torch.distributed as dist
    import zkvm_prover_b200 as z
    from zkvm_prover_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = z.Context(rank)
        full = ctx.alloc(n, w).fill(4321).to_host()
        wg = w // world
        local = np.ascontiguousarray(full[:, rank * wg:(rank + 1) * wg])
        root, cap = D.sharded_lde_commit(D.GpuOps(ctx), local, 1, z.GENERATOR_MONTY)
        np.save(os.path.join(out_dir, f"root{rank}.npy"), root)
        # the same with the exchange fused into the last NTT pass (TMA stores into the peer's receive buffer), twice on one mapping
        exch = D.PeerExchange(ctx, 2 * n, wg)
        for rep in range(4):  # both receive layouts (one wide matrix per owner / one slot per sender), twice each on one mapping
            root2, cap2 = D.sharded_lde_commit_p2p(ctx, ctx.upload(local), 1, z.GENERATOR_MONTY, exch, rows_layout=rep % 2 == 0)
            assert np.array_equal(cap2, cap), (rank, rep)
            np.save(os.path.join(out_dir, f"p2p{rank}_{rep % 2}.npy"), root2)
        exch.close()
        if rank == 0:  # single-GPU reference on the same device
            pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=1), ctx)
            r1, _ = pcs.commit([full])
            np.save(os.path.join(out_dir, "single.npy"), r1)
    finally:
        dist.destroy_process_group()


def test_sharded_commit_on_gpus(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), 1 << 14, 128, str(tmp_path)), nprocs=world, join=True)   # 64 columns per rank: 4 pipelined strips
    single = np.load(tmp_path / "single.npy")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"root{r}.npy"), single)
        assert np.array_equal(np.load(tmp_path / f"p2p{r}_0.npy"), single)
        assert np.array_equal(np.load(tmp_path / f"p2p{r}_1.npy"), single)
