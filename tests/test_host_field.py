"""Host-side field glue of the mirror (zkvm_prover_b200/field.py) against the pure-Python oracle: representation
conversions and the few EF4 helpers the open phase uses on the host.  CPU only."""
import importlib.util
import os

import numpy as np

from oracle import pyref as R

_spec = importlib.util.spec_from_file_location("b200zk_field", os.path.join(os.path.dirname(__file__), "..", "zkvm_prover_b200", "field.py"))
F = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(F)
P = R.P


def _ef(rng):
    return F.to_monty(rng.integers(0, P, 4, dtype=np.uint64))


def _canon(a):
    return [int(x) for x in F.from_monty(a)]


def test_montgomery_conversions_and_generators():
    rng = np.random.default_rng(1)
    x = rng.integers(0, P, 1000, dtype=np.uint64)
    m = F.to_monty(x)
    assert [int(v) for v in m[:50]] == [R.to_monty(int(v)) for v in x[:50]]
    assert np.array_equal(F.from_monty(m), x.astype(np.uint32))
    assert F.MONTY_ONE == R.to_monty(1) and F.GENERATOR_MONTY == R.to_monty(31)
    assert F.monty_scalar(P + 5) == R.to_monty(5)
    for bits in range(28):
        assert F.two_adic_generator(bits) == R.two_adic_generator(bits)
    g = F.two_adic_generator(27)
    assert pow(g, 1 << 26, P) == P - 1          # order exactly 2^27


def test_ef4_helpers_match_the_oracle():
    rng = np.random.default_rng(2)
    a, b = _ef(rng), _ef(rng)
    assert _canon(F.ef_mul(a, b)) == R.ef_mul(_canon(a), _canon(b))
    assert _canon(F.ef_add(a, b)) == R.ef_add(_canon(a), _canon(b))
    assert _canon(F.ef_scale_base(a, 12345)) == R.ef_scale(_canon(a), 12345)
    assert _canon(F.ef_pow(a, 77)) == R.ef_pow(_canon(a), 77)
    assert _canon(F.ef_pow(a, 0)) == [1, 0, 0, 0]
    for n in (0, 1, 2, 3, 17, 64, 100):
        pw = F.ef_powers(a, n)
        assert pw.shape == (n, 4)
        acc = [1, 0, 0, 0]
        for i in range(n):
            assert _canon(pw[i]) == acc, (n, i)
            acc = R.ef_mul(acc, _canon(a))
    ys = F.to_monty(rng.integers(0, P, (37, 4), dtype=np.uint64))
    pw = F.ef_powers(a, 37)
    acc = [0, 0, 0, 0]
    for i in range(37):
        acc = R.ef_add(acc, R.ef_mul(_canon(pw[i]), _canon(ys[i])))
    assert _canon(F.ef_dot(pw, ys)) == acc
    assert _canon(F.ef_dot(pw[:0], ys[:0])) == [0, 0, 0, 0]
    # extreme values: every coordinate p - 1
    top = F.to_monty(np.full((5, 4), P - 1, dtype=np.uint64))
    exp = [0, 0, 0, 0]
    for i in range(5):
        exp = R.ef_add(exp, R.ef_mul([P - 1] * 4, [P - 1] * 4))
    assert _canon(F.ef_dot(top, top)) == exp
