"""BabyBear host helpers: representation conversions only (numpy), no prover arithmetic.

p = 2^31 - 2^27 + 1; elements cross the C ABI in Montgomery form (R = 2^32), the in-memory bytes of
p3_baby_bear::BabyBear (p3-monty-31 0.4.3, Cargo.lock:5685 of the reference).
"""
from __future__ import annotations

import numpy as np

P = 0x78000001
MONTY_ONE = 0x0FFFFFFE
_RINV = pow(1 << 32, -1, P)


def to_monty(x):
    """canonical integers -> Montgomery-form uint32 array"""
    x = np.asarray(x, dtype=np.uint64) % np.uint64(P)
    return ((x << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def from_monty(m):
    """Montgomery-form uint32 -> canonical uint32 (BabyBear::as_canonical_u32)"""
    m = np.asarray(m, dtype=np.uint64)
    return ((m * np.uint64(_RINV)) % np.uint64(P)).astype(np.uint32)


def monty_scalar(x: int) -> int:
    return (int(x) % P) * (1 << 32) % P


def two_adic_generator(bits: int) -> int:
    """BabyBear::two_adic_generator(bits), canonical integer (31^15 generates the 2^27 subgroup)."""
    if not 0 <= bits <= 27:
        raise ValueError("BabyBear has two-adicity 27")
    return pow(pow(31, 15, P), 1 << (27 - bits), P)


GENERATOR_MONTY = monty_scalar(31)  # BabyBear::GENERATOR, the coset shift of trace-domain LDEs


# ---- EF4 = F[x]/(x^4 - 11) scalar helpers for host-side glue (a handful of elements per proof; Montgomery u32 in/out)
def _c(a):
    return [int(x) * _RINV % P for x in np.asarray(a, dtype=np.uint64).reshape(-1)]


def _m(a):
    return np.array([(x % P) * (1 << 32) % P for x in a], dtype=np.uint32)


def ef_mul(a, b):
    a, b = _c(a), _c(b)
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] += a[i] * b[j]
    return _m([t[i] + 11 * (t[i + 4] if i < 3 else 0) for i in range(4)])


def ef_add(a, b):
    return _m([x + y for x, y in zip(_c(a), _c(b))])


def ef_scale_base(a, k_canonical: int):
    return _m([x * k_canonical for x in _c(a)])


def ef_pow(a, e: int):
    r = _m([1, 0, 0, 0])
    b = np.asarray(a, dtype=np.uint32)
    while e:
        if e & 1:
            r = ef_mul(r, b)
        b = ef_mul(b, b)
        e >>= 1
    return r


# ---- vectorised EF4 helpers (numpy, canonical uint64 inside): the per-column glue of the open phase
def _ef_mul_canon(a, b):
    """a, b: (..., 4) canonical uint64 arrays -> (..., 4) canonical"""
    p = np.uint64(P)
    pr = [[(a[..., i] * b[..., j]) % p for j in range(4)] for i in range(4)]   # 31-bit x 31-bit fits in 64 bits
    w = np.uint64(11)
    c0 = (pr[0][0] + w * ((pr[1][3] + pr[2][2] + pr[3][1]) % p)) % p
    c1 = (pr[0][1] + pr[1][0] + w * ((pr[2][3] + pr[3][2]) % p)) % p
    c2 = (pr[0][2] + pr[1][1] + pr[2][0] + w * pr[3][3]) % p
    c3 = (pr[0][3] + pr[1][2] + pr[2][1] + pr[3][0]) % p
    return np.stack([c0, c1, c2, c3], axis=-1)


def _canon_arr(a):
    return (np.asarray(a, dtype=np.uint64) * np.uint64(_RINV)) % np.uint64(P)


def ef_powers(a, n: int):
    """[a^0, a^1, ..., a^(n-1)] as an (n, 4) Montgomery array (doubling: log2 n vectorised products)"""
    base = _canon_arr(a).reshape(4)
    pw = np.zeros((1, 4), np.uint64)
    pw[0, 0] = 1
    step = base.copy()                       # a^len(pw)
    while pw.shape[0] < n:
        pw = np.concatenate([pw, _ef_mul_canon(pw, step[None, :])], axis=0)
        step = _ef_mul_canon(step, step)
    return to_monty(pw[:n])


def ef_dot(a, b):
    """sum_i a[i] * b[i] for (n, 4) Montgomery arrays -> EF4 (Montgomery)"""
    a, b = _canon_arr(a).reshape(-1, 4), _canon_arr(b).reshape(-1, 4)
    if a.shape[0] == 0:
        return np.zeros(4, np.uint32)
    prod = _ef_mul_canon(a, b)
    return to_monty(prod.sum(axis=0) % np.uint64(P))   # at most 2^33 terms of 31 bits would fit; widths are << that
