"""Host mirror of the p3_symmetric objects of the reference's hash configuration
(openvm_stark_sdk::config::baby_bear_poseidon2): Permutation, CryptographicHasher,
PseudoCompressionFunction -- each call runs the CUDA kernels (batched variants for throughput)."""
from __future__ import annotations

import numpy as np

from .device import Context, DeviceMatrix, default_context

WIDTH, RATE, DIGEST = 16, 8, 8


class Poseidon2BabyBear16:
    """Permutation<[BabyBear; 16]>: default_perm() of the reference (Horizen RC16, x^7, 4+13+4)."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    def permute(self, state):
        s = np.ascontiguousarray(state, dtype=np.uint32).copy()
        if s.size % WIDTH:
            raise ValueError("state width must be 16")
        self.ctx.check(self.ctx.lib.b200zk_poseidon2_permute(self.ctx.h, s.ctypes.data, s.size // WIDTH))
        return s

    permute_mut = permute


class PaddingFreeSponge:
    """CryptographicHasher<BabyBear, [BabyBear; 8]> = PaddingFreeSponge<Perm, 16, 8, 8>."""

    def __init__(self, perm: Poseidon2BabyBear16 | None = None):
        self.perm = perm or Poseidon2BabyBear16()
        self.ctx = self.perm.ctx

    def hash_iter(self, items):
        return self.hash_slice(np.fromiter(items, dtype=np.uint32))

    def hash_slice(self, items):
        v = np.ascontiguousarray(items, dtype=np.uint32).reshape(1, -1)
        if v.size == 0:
            return np.zeros(DIGEST, np.uint32)  # hash_iter of nothing returns the zero state prefix
        return self.hash_rows(v)[0]

    def hash_iter_slices(self, slices):
        return self.hash_slice(np.concatenate([np.asarray(s, dtype=np.uint32).reshape(-1) for s in slices]))

    def hash_item(self, item):
        return self.hash_slice([item])

    def hash_rows(self, mat):
        """batched: one digest per row of a matrix (host array or DeviceMatrix)"""
        m = mat if isinstance(mat, DeviceMatrix) else self.ctx.upload(mat)
        out = np.empty((m.rows, DIGEST), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_hash_rows(self.ctx.h, m.h, out.ctypes.data))
        return out


class TruncatedPermutation:
    """PseudoCompressionFunction<[BabyBear; 8], 2> = TruncatedPermutation<Perm, 2, 8, 16>."""

    def __init__(self, perm: Poseidon2BabyBear16 | None = None):
        self.perm = perm or Poseidon2BabyBear16()
        self.ctx = self.perm.ctx

    def compress(self, pair):
        p = np.ascontiguousarray(pair, dtype=np.uint32).reshape(-1, 2 * DIGEST)
        out = np.empty((p.shape[0], DIGEST), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_compress_pairs(self.ctx.h, p.ctypes.data, out.ctypes.data, p.shape[0]))
        return out[0] if out.shape[0] == 1 and np.ndim(pair) <= 2 and np.shape(pair)[0] == 2 else out
