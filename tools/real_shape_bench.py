"""Times TwoAdicFriPcs::commit on the real 17-AIR shape of the reference fixture (see tests/golden/real_shape_commit.json)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "real_shape_commit.json")))
ctx = z.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
traces = [ctx.alloc(d, w).fill(g["seed_base"] + i) for i, (d, w) in enumerate(zip(g["degrees"], g["widths"]))]
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=g["log_blowup"]), ctx)
dft = z.B200Dft(ctx)
for _ in range(2):
    root, pd = pcs.commit(traces); pd.free()
ctx.sync()
times = []
for _ in range(7):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record(stream)
    root, pd = pcs.commit(traces); pd.free()
    e[1].record(stream); ctx.sync(); torch.cuda.synchronize()
    times.append(e[0].elapsed_time(e[1]))
elems = sum(d * w for d, w in zip(g["degrees"], g["widths"]))
print(f"real shape commit (17 AIRs, {elems/1e6:.1f} M trace elements, blowup 4): min {min(times):.2f} ms, median {sorted(times)[3]:.2f} ms  root ok: {root.tolist() == g['root']}")
# per-matrix LDE time for the big ones
for i, (d, w) in enumerate(zip(g["degrees"], g["widths"])):
    if d * w < (1 << 21): continue
    out = ctx.alloc(d << 2, w)
    dft.coset_lde_batch(traces[i], 2, z.GENERATOR_MONTY, bit_reversed=True, out=out); ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); dft.coset_lde_batch(traces[i], 2, z.GENERATOR_MONTY, bit_reversed=True, out=out); b.record(stream); ctx.sync(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"  LDE 2^{d.bit_length()-1} x {w:3d}: {ms:7.3f} ms  {4*d*w*5/ms/1e6:7.1f} GB/s algorithmic")
    out.free()
