"""b200zk_lde_commit_host_async as a stream (two commits in flight): host-side time of every issue / collect / free call and the
steady-state time per commit (pinned 2^23 x 256 trace)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zkvm_prover_b200 as z
ctx = z.default_context(0)
pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=1), ctx)
shape = (1 << int(os.environ.get("LOG_ROWS", "23")), 256)
host = torch.empty(shape, dtype=torch.int32).pin_memory()
host.random_(0, 2013265921)
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
for i, (t, a, b, c) in enumerate(marks):
    print(f"commit {i}: at {1e3 * t:8.1f} ms   issue {1e3 * a:7.1f}  collect(prev) {1e3 * b:7.1f}  free(prev) {1e3 * c:6.1f}")
print(f"B200ZK_STRIP={strip}: {K} commits in {1e3 * (t_end - t_start):.1f} ms; steady state {(marks[-1][0] - marks[2][0]) * 1e3 / (K - 3):.1f} ms per commit; root {root[:2].tolist()}")
