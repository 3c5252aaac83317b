"""Phase timing of one real-shape segment (commit, open-phase pieces, FRI commit phase, PoW, queries); host wall clock with syncs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkvm_prover_b200 as z
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "real_shape_commit.json")))
ctx = z.Context(0)
cfg = z.FriConfig(log_blowup=g["log_blowup"], log_final_poly_len=0, num_queries=100, proof_of_work_bits=16)
pcs = z.TwoAdicFriPcs(cfg, ctx)
traces = [ctx.alloc(d, w).fill(g["seed_base"] + i) for i, (d, w) in enumerate(zip(g["degrees"], g["widths"]))]
import zkvm_prover_b200.fri as F
T = {}
def timed(name, fn):
    def w(*a, **k):
        ctx.sync(); t = time.perf_counter(); r = fn(*a, **k); ctx.sync(); T[name] = T.get(name, 0) + time.perf_counter() - t; return r
    return w
for name in ["dot_ext_powers", "inv_denominators", "interpolate_coset", "reduce_openings"]:
    setattr(pcs, name, timed(name, getattr(pcs, name)))
F.commit_phase = timed("fri_commit_phase", F.commit_phase)
pcs.mmcs.open_batch_many = timed("open_batch_many", pcs.mmcs.open_batch_many)
for it in range(3):
    T.clear()
    ctx.sync(); t0 = time.perf_counter()
    root, pd = pcs.commit(traces); ctx.sync(); t1 = time.perf_counter()
    ch = z.DuplexChallenger(ctx); ch.observe(root); zeta = ch.sample_algebra_element()
    pts = [[zeta, z.field.ef_scale_base(zeta, z.field.two_adic_generator(d.bit_length() - 1))] for d in g["degrees"]]
    ch.grind = timed("grind", ch.grind)
    opened, proof = pcs.open([(pd, pts)], ch); ctx.sync(); t2 = time.perf_counter()
    print(f"iter {it}: commit {1e3*(t1-t0):.1f} ms, open {1e3*(t2-t1):.1f} ms  ::", {k: round(1e3 * v, 1) for k, v in T.items()})
    pd.free()
