"""CPU: the product's device arithmetic headers compiled for the host (tests/host_check.cpp) must agree
bit-for-bit with the oracle.  Guards the instruction-saving formulations used in the CUDA kernels."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = O.P


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hc") / "host_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-x", "c++", os.path.join(ROOT, "tests", "host_check.cpp"), "-o", out], check=True)
    return out


def run(exe, op, arr):
    r = subprocess.run([exe, op], input=np.ascontiguousarray(arr, dtype=np.uint32).tobytes(), capture_output=True, check=True)
    return np.frombuffer(r.stdout, dtype=np.uint32)


def edge_states():
    rng = np.random.default_rng(11)
    st = rng.integers(0, P, (4000, 16), dtype=np.uint64).astype(np.uint32)
    st[0] = 0
    st[1] = P - 1
    st[2] = np.arange(16)
    st[3] = O.to_monty(np.arange(16))
    st[4, ::2] = P - 1
    st[5, 1::2] = 1
    return st


@pytest.mark.parametrize("op", ["permute", "permute_plain"])
def test_host_permute_matches_oracle(exe, op):
    st = edge_states()
    got = run(exe, op, st).reshape(-1, 16)
    assert np.array_equal(got, O.permute(st))


def test_host_field_helpers(exe):
    rng = np.random.default_rng(12)
    x = rng.integers(0, P, 3000, dtype=np.uint64).astype(np.uint32)
    x[:4] = [0, 1, P - 1, P - 2]
    got = run(exe, "div2exp", x).reshape(-1, 6)
    L = O.lib()
    for j, k in enumerate([1, 2, 3, 4, 8, 27]):
        c = int(O.to_monty([pow(2, -k, P)])[0])
        exp = np.array([L.orc_mul(int(v), c) for v in x[:300]], dtype=np.uint32)
        assert np.array_equal(got[:300, j], exp), k
    ab = rng.integers(0, P, (500, 2), dtype=np.uint64).astype(np.uint32)
    ab[0] = [0, 0]
    ab[1] = [P - 1, P - 1]
    ab[2] = [P - 1, 0]
    got = run(exe, "mul", ab).reshape(-1, 3)
    for (a, b), (m, sm, g) in zip(ab, got):
        assert m == L.orc_mul(int(a), int(b)) and sm == m
        assert g == L.orc_two_adic_generator(int(a) % 28)
    e = rng.integers(0, P, (300, 8), dtype=np.uint64).astype(np.uint32)
    got = run(exe, "efmul", e).reshape(-1, 4)
    for row, g in zip(e, got):
        out = np.zeros(4, np.uint32)
        L.orc_ef_mul(np.ascontiguousarray(row[:4]), np.ascontiguousarray(row[4:]), out)
        assert np.array_equal(out, g)
