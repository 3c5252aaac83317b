"""cProfile of one real-shape segment's open() (host-side glue hot spots)."""
import cProfile, json, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkvm_prover_b200 as z
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "real_shape_commit.json")))
ctx = z.Context(0)
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=g["log_blowup"], log_final_poly_len=0, num_queries=100, proof_of_work_bits=16), ctx)
traces = [ctx.alloc(d, w).fill(g["seed_base"] + i) for i, (d, w) in enumerate(zip(g["degrees"], g["widths"]))]
def seg():
    root, pd = pcs.commit(traces)
    ch = z.DuplexChallenger(ctx); ch.observe(root); zeta = ch.sample_algebra_element()
    pts = [[zeta, z.field.ef_scale_base(zeta, z.field.two_adic_generator(d.bit_length() - 1))] for d in g["degrees"]]
    return pcs.open([(pd, pts)], ch)
for _ in range(3): seg()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): seg()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
