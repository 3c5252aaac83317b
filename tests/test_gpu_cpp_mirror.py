"""GPU: the C++ host mirror (include/b200zk.hpp) builds against libb200zk.so and passes its checks."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_mirror_runs(tmp_path):
    exe = str(tmp_path / "cpp_mirror_check")
    libdir = os.path.join(ROOT, "zkvm_prover_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                    "-L", libdir, "-lb200zk", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout

    # the C++ open phase (TwoAdicFriPcs::open + FriProof::encode) against the Python mirror on the same deterministic input:
    # identical transcript => identical proof bytes and opened values
    import numpy as np
    import zkvm_prover_b200 as z
    from zkvm_prover_b200.proof import FriProof
    pf, of = str(tmp_path / "proof.bin"), str(tmp_path / "opened.bin")
    r = subprocess.run([exe, pf, of], capture_output=True, text=True)
    assert r.returncode == 0 and "cpp open ok" in r.stdout, r.stdout + r.stderr
    ctx = z.default_context(0)
    pcs = z.TwoAdicFriPcs(z.FriConfig(log_blowup=1, log_final_poly_len=0, num_queries=5, proof_of_work_bits=6), ctx)

    def gen(n, w, seed):
        i = np.arange(n * w, dtype=np.uint64)
        return z.to_monty((i * np.uint64(2654435761) + np.uint64(17 + seed)) % np.uint64(z.P)).reshape(n, w)

    (root1, pd1), (root2, pd2) = pcs.commit([gen(1 << 9, 12, 1), gen(1 << 7, 5, 2)]), pcs.commit([gen(1 << 9, 3, 3)])
    ch = z.DuplexChallenger(ctx)
    ch.observe(root1)
    ch.observe(root2)
    zeta = ch.sample_algebra_element()
    pts = lambda ln: [zeta, z.field.ef_scale_base(zeta, z.field.two_adic_generator(ln))]  # noqa: E731
    opened, proof = pcs.open([(pd1, [pts(9), pts(7)]), (pd2, [pts(9)])], ch)
    assert FriProof.from_pcs_open(proof).encode() == open(pf, "rb").read()
    flat = np.concatenate([y.reshape(-1) for rr in opened for m in rr for y in m])
    assert np.array_equal(flat, np.fromfile(of, dtype=np.uint32))
