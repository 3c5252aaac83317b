"""Times the open-phase kernels on the committed 2^24 x 256 LDE (HBM-bound part of the path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkvm_prover_b200 as z
n, w, b = (int(sys.argv[1]), int(sys.argv[2]), 1) if len(sys.argv) > 2 else (23, 256, 1)
ctx = z.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=b), ctx)
tr = ctx.alloc(1 << n, w).fill(5)
root, pd = pcs.commit([tr])
lde = pd.mats[0]
m = lde.rows
alpha = np.array([11, 22, 33, 44], np.uint32); zp = np.array([5, 6, 7, 8], np.uint32)
def timed(fn, reps=5):
    fn(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): r = fn()
    e1.record(stream); ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
peak = 6551.0
ms, inv = timed(lambda: pcs.inv_denominators(n + b, zp))
print(f"denominators 2^{n+b}: {ms:.3f} ms ({16*m/ms/1e6:.0f} GB/s written)")
ms, rr = timed(lambda: pcs.dot_ext_powers(lde, alpha))
byt = 4 * m * w + 16 * m
print(f"dot_ext_powers 2^{n+b} x {w}: {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s = {byt/ms/1e6/peak:.2%} of HBM roofline")
ms, ys = timed(lambda: pcs.interpolate_coset(lde, zp, inv))
byt = 4 * (m >> b) * w + 16 * (m >> b)
print(f"interpolate_coset (low coset 2^{n} x {w}): {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s = {byt/ms/1e6/peak:.2%} of HBM roofline")
ro = z.DeviceBuffer(ctx, 16 * m).zero()
ms, _ = timed(lambda: pcs.reduce_openings(rr, m, inv, alpha, zp, ro))
byt = 64 * m
print(f"reduce_openings 2^{n+b}: {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s = {byt/ms/1e6/peak:.2%} of HBM roofline")
