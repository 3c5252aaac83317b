"""Times the coset LDE (and optionally the commit) for a few shapes; prints ms, algorithmic GB/s and an output checksum
(so kernel variants selected with B200ZK_LIB_PATH can be compared for speed and for bit-identical results)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z

shapes = [(20, 64), (22, 256), (23, 256)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
ctx = z.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
dft = z.B200Dft(ctx)
tag = os.path.basename(os.environ.get("B200ZK_LIB_PATH", "default"))
for n, w in shapes:
    tr = ctx.alloc(1 << n, w).fill(7)
    out = ctx.alloc(2 << n, w)
    for _ in range(3):
        dft.coset_lde_batch(tr, 1, z.GENERATOR_MONTY, bit_reversed=True, out=out)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        dft.coset_lde_batch(tr, 1, z.GENERATOR_MONTY, bit_reversed=True, out=out)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    b = 4 * (1 << n) * w * 3
    print(f"{tag:16s} LDE 2^{n}x{w}: {ms:8.3f} ms  {b / ms / 1e6:8.1f} GB/s algorithmic  checksum {out.checksum():016x}", flush=True)
    tr.free(); out.free()
