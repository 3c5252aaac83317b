"""Small driver for ncu captures: runs each hot kernel once or twice at a moderate size."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_prover_b200 as z

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
width = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = z.default_context(0)
lib = ctx.lib
tr = ctx.alloc(1 << log_rows, width).fill(1)
dft = z.B200Dft(ctx)
lde = ctx.alloc(2 << log_rows, width)
for _ in range(2):
    dft.coset_lde_batch(tr, 1, z.GENERATOR_MONTY, bit_reversed=True, out=lde)
mm = z.MerkleTreeMmcs(ctx)
for _ in range(2):
    root, pd = mm.commit([lde])
    pd.mats = []
    pd.free()
st = ctx.alloc(1 << 22, 16).fill(2)
for _ in range(2):
    ctx.check(lib.b200zk_poseidon2_permute_dev(ctx.h, st.device_ptr, 1 << 22))
ctx.sync()
print("done", root.tolist())
