"""Strong scaling of ONE wide matrix (second row): 2^log_rows x width trace, columns sharded over the ranks for the
LDE, all-to-all (NCCL over NVLink) to row blocks, per-rank subtree, cap gather.  Correctness of the scheme (bit-identical root) is
tests/test_gpu_dist.py + tests/test_dist_cpu.py; this tool only times it.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_bench.py [log_rows width]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import zkvm_prover_b200 as z
from zkvm_prover_b200 import dist as D

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 23
width = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = z.Context(local)
ops = D.GpuOps(ctx)
wg = width // world
shard = torch.empty((1 << log_rows, wg), dtype=torch.int32, device="cuda")
m = ctx.wrap(shard.data_ptr(), 1 << log_rows, wg, keepalive=shard)
m.fill(1000 + rank)
ctx.sync()
T = {}
if os.environ.get("SHARD_PHASES"):   # phase timing with host syncs (serialises the overlap: for diagnosis only)
    def timed(name, fn):
        def w(*a, **k):
            torch.cuda.synchronize(); ctx.sync(); t = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize(); ctx.sync()
            T[name] = T.get(name, 0) + time.perf_counter() - t
            return r
        return w
    ops.lde = timed("lde", ops.lde)
    ops.subtree_root = timed("subtree_root", ops.subtree_root)
    D._all_to_all_start = timed("all_to_all", lambda recv, send, group=None: (D._all_to_all_equal(recv, send, group), None)[1])
p2p = bool(os.environ.get("SHARD_P2P"))
exch = D.PeerExchange(ctx, 2 << log_rows, wg) if p2p else None
times = []
for it in range(5):
    T.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    root, cap = D.sharded_lde_commit_p2p(ctx, m, 1, z.GENERATOR_MONTY, exch) if p2p else D.sharded_lde_commit(ops, shard, 1, z.GENERATOR_MONTY)
    ctx.sync(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt.item()))
if rank == 0:
    best = min(times[2:])
    gb = 4 * (1 << log_rows) * width * 3 + 4 * (2 << log_rows) * width + 32 * ((4 << log_rows) - 1)
    print(json.dumps({"workload": f"one matrix 2^{log_rows} x {width}, column-sharded LDE + " + ("TMA stores into peer memory" if p2p else "NCCL all-to-all") + " + subtree commit", "n_gpus": world,
                      "ms": round(1e3 * best, 2), "GB_per_s": round(gb / best / 1e9, 1), "all_ms": [round(1e3 * t, 1) for t in times],
                      "root": [int(x) for x in root], "phases_ms": {k: round(1e3 * v, 1) for k, v in T.items()}}))
if exch:
    exch.close()
dist.destroy_process_group()
