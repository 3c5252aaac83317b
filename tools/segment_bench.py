"""BASELINE config [3]: a chunk proof's continuation segments sharded across GPUs.

Every segment is one STARK-shaped PCS job on the REAL 17-AIR shape of the reference's aggregation-layer fixture
(tests/golden/real_shape_commit.json): commit of the main traces (coset LDE + MMCS), sample zeta, open phase
(opened values at zeta and zeta*g, reduced openings), FRI commit phase, PoW grinding, query openings.  Segments are
independent: rank r takes segments r, r + world, ... (zkvm_prover_b200.dist.segment_assignment); the only collective is
the all-gather of the segment commitments at the end.  Trace generation / constraint evaluation are outside the path.

    python tools/segment_bench.py --segments 8
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/segment_bench.py --segments 8
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import zkvm_prover_b200 as z
from zkvm_prover_b200.dist import segment_assignment

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=8)
ap.add_argument("--queries", type=int, default=100)
ap.add_argument("--pow-bits", type=int, default=16)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "real_shape_commit.json")))
ctx = z.Context(local)
cfg = z.FriConfig(log_blowup=g["log_blowup"], log_final_poly_len=0, num_queries=args.queries, proof_of_work_bits=args.pow_bits)
pcs = z.TwoAdicFriPcs(cfg, ctx)
mine = segment_assignment(args.segments, world)[rank]

def prove_segment(seg, traces):
    root, pd = pcs.commit(traces)
    ch = z.DuplexChallenger(ctx)
    ch.observe(root)
    zeta = ch.sample_algebra_element()
    pts = [[zeta, z.field.ef_scale_base(zeta, z.field.two_adic_generator(d.bit_length() - 1))] for d in g["degrees"]]
    opened, proof = pcs.open([(pd, pts)], ch)
    return root, proof

traces = [ctx.alloc(d, w).fill(g["seed_base"] + i) for i, (d, w) in enumerate(zip(g["degrees"], g["widths"]))]
prove_segment(-1, traces)  # warm-up (twiddle caches, memory pool)
ctx.sync(); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
roots = []
for seg in mine:
    for i, t in enumerate(traces):  # a different witness per segment
        t.fill(g["seed_base"] + 1000 * (seg + 1) + i)
    ts = time.perf_counter()
    root, proof = prove_segment(seg, traces)
    roots.append(root)
    if os.environ.get("SEGMENT_TIMES"):
        ctx.sync()
        print(f"rank {rank} segment {seg}: {1e3 * (time.perf_counter() - ts):.1f} ms", flush=True)
ctx.sync(); torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    allr = [None] * world
    dist.all_gather_object(allr, [r.tolist() for r in roots])   # the only collective: gather the segment commitments
    n_roots = sum(len(x) for x in allr)
else:
    n_roots = len(roots)
if rank == 0:
    el = sum(d * w for d, w in zip(g["degrees"], g["widths"]))
    print(json.dumps({"workload": "segment-sharded PCS (commit + open + FRI commit phase + PoW + queries) on the real 17-AIR shape",
                      "segments": args.segments, "n_gpus": world, "seconds": round(float(dt.item()), 4),
                      "segments_per_s": round(args.segments / float(dt.item()), 3), "trace_elements_per_segment": el,
                      "queries": args.queries, "pow_bits": args.pow_bits, "roots_gathered": n_roots}))
if world > 1:
    dist.destroy_process_group()
