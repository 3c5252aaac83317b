"""e2e timing of b200zk_lde_commit_host (pinned 2^23 x 256 trace), min of 5 calls; env knobs of the library apply."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z
ctx = z.default_context(0)
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=1), ctx)
host = torch.empty((1 << 23, 256), dtype=torch.int32).pin_memory()
host.random_(0, 2013265921)
ts = []
for i in range(6):
    t0 = time.perf_counter()
    root, pd = pcs.commit_host(None, strip_cols=int(os.environ.get("B200ZK_STRIP", "0")), host_ptr=host.data_ptr(), shape=(1 << 23, 256))
    ts.append(time.perf_counter() - t0)
    pd.free()
print(f"B200ZK_STRIP={os.environ.get('B200ZK_STRIP', '0 (library schedule)')}: e2e min {1e3 * min(ts[1:]):.1f} ms  all {[round(1e3 * t, 1) for t in ts]}  root {root[:2].tolist()}")
