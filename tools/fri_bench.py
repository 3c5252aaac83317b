"""Times the whole FRI commit phase on the device (commit -> observe -> sample beta -> fold, repeated), SURVEY 8(d):
2^25 EF4 elements, folded down to blowup*final_poly_len = 2, challenger on the device."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zkvm_prover_b200 as z
ln = int(sys.argv[1]) if len(sys.argv) > 1 else 25
ctx = z.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
vec = ctx.alloc(1 << ln, 4).fill(99)       # 2^ln EF4 elements (4 base coefficients each)
chal = z.DuplexChallenger(ctx)
chal.observe(np.arange(8, dtype=np.uint32))
cfg = z.FriConfig(log_blowup=1, log_final_poly_len=0)
def run():
    return z.commit_phase(cfg, [(vec.device_ptr, 1 << ln)], chal, ctx, keep_trees=False)
for _ in range(2): r = run()
ctx.sync()
l0 = ctx.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record(stream)
for _ in range(reps): r = run()
e1.record(stream); ctx.sync(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byt = sum(16 * (1 << (ln - k)) * 1.5 + 32 * (1 << (ln - k)) for k in range(len(r.commits)))
print(f"FRI commit phase 2^{ln} EF4: {len(r.commits)} rounds, {ms:.2f} ms, {(ctx.launches - l0)//reps} kernels/phase, algorithmic {byt/1e9:.2f} GB -> {byt/ms/1e6:.1f} GB/s; first root {r.commits[0][:3].tolist()}")
