"""b200zk_lde_commit_host_async as a stream (two commits in flight): host-side time of every issue / collect / free call and the
steady-state time per commit (pinned 2^23 x 256 trace)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z
import ctypes as C
import numpy as np
local = int(os.environ.get("LOCAL_RANK", "0"))     # under torchrun: one process per GPU, all streaming at once (no collective needed)
torch.cuda.set_device(local)
ctx = z.default_context(local)
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=1), ctx)
shape = (1 << int(os.environ.get("LOG_ROWS", "23")), 256)
wc = int(os.environ.get("HOST_WC", "0"))           # 1: write-combined pinned memory from b200zk_host_alloc
hp = C.c_void_p()
ctx.check(ctx.lib.b200zk_host_alloc(4 * shape[0] * shape[1], wc, C.byref(hp)))
harr = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=shape)
rng = np.random.default_rng(local)
blk = rng.integers(0, 2013265921, (1 << 16, shape[1]), dtype=np.uint64).astype(np.uint32)
for r0 in range(0, shape[0], 1 << 16):               # streaming stores only (write-combined pages must not be read back)
    harr[r0:r0 + (1 << 16)] = blk


class _Host:
    def data_ptr(self):
        return hp.value


host = _Host()
strip = int(os.environ.get("B200ZK_STRIP", "0"))
K = 8
pend = None
t_start = time.perf_counter()
marks = []
for i in range(K):
    t0 = time.perf_counter()
    p = pcs.commit_host_async(host.data_ptr(), shape, strip_cols=strip)
    t1 = time.perf_counter()
    if pend is not None:
        root, pd = pend.result()
        t2 = time.perf_counter()
        pd.free()
    else:
        t2 = t1
    t3 = time.perf_counter()
    marks.append((t3 - t_start, t1 - t0, t2 - t1, t3 - t2))
    pend = p
root, pd = pend.result()
pd.free()
t_end = time.perf_counter()
for i, (t, a, b, c) in enumerate(marks if local == 0 and not os.environ.get("QUIET") else []):
    print(f"commit {i}: at {1e3 * t:8.1f} ms   issue {1e3 * a:7.1f}  collect(prev) {1e3 * b:7.1f}  free(prev) {1e3 * c:6.1f}")
print(f"rank {local} HOST_WC={wc} B200ZK_STRIP={strip}: {K} commits in {1e3 * (t_end - t_start):.1f} ms; steady state {(marks[-1][0] - marks[2][0]) * 1e3 / (K - 3):.1f} ms per commit; root {root[:2].tolist()}")
