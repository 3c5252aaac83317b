"""Times the leaf-hash kernel (b200zk_hash_rows_dev) on a resident rows x width matrix; B200ZK_LIB_PATH selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24
width = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = z.default_context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
m = ctx.alloc(1 << log_rows, width).fill(3)
dig = z.DeviceBuffer(ctx, (1 << log_rows) * 32)
for _ in range(2):
    ctx.check(ctx.lib.b200zk_hash_rows_dev(ctx.h, m.h, dig.ptr))
ctx.sync()
ts = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    ctx.check(ctx.lib.b200zk_hash_rows_dev(ctx.h, m.h, dig.ptr))
    b.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
perms = (1 << log_rows) * ((width + 7) // 8)
import numpy as np
d = dig.to_host((4, 8))
print(f"{os.environ.get('B200ZK_LIB_PATH', 'default'):40s} leaf hash 2^{log_rows} x {width}: min {min(ts):.3f} ms  {perms / min(ts) / 1e6:.3f} Gperm/s  digest0 {d[0, :2].tolist()}")
