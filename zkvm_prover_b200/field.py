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
