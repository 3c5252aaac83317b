"""Host mirror of p3_dft::TwoAdicSubgroupDft (p3-dft 0.4.3) backed by the CUDA NTT.

Method names, argument meaning and panics (-> exceptions) follow the trait: matrices are
RowMajorMatrix<BabyBear> (numpy uint32, Montgomery form, or DeviceMatrix), `shift` is a Montgomery
field element, heights must be powers of two.  Methods return DeviceMatrix (device resident);
`.to_host()` materialises the RowMajorMatrix.
"""
from __future__ import annotations

import ctypes as C

from .device import Context, DeviceMatrix, default_context
from .field import MONTY_ONE


class B200Dft:
    """Drop-in for `Radix2DitParallel<BabyBear>` as the `Dft` of TwoAdicFriPcs."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    def _dev(self, mat) -> DeviceMatrix:
        return mat if isinstance(mat, DeviceMatrix) else self.ctx.upload(mat)

    def _dft(self, mat, shift, inverse, bitrev) -> DeviceMatrix:
        m = self._dev(mat)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_dft_batch(self.ctx.h, m.h, shift, int(inverse), int(bitrev), C.byref(h)))
        return DeviceMatrix(self.ctx, h, True)

    def dft_batch(self, mat) -> DeviceMatrix:
        return self._dft(mat, MONTY_ONE, False, False)

    def coset_dft_batch(self, mat, shift: int) -> DeviceMatrix:
        return self._dft(mat, shift, False, False)

    def idft_batch(self, mat) -> DeviceMatrix:
        return self._dft(mat, MONTY_ONE, True, False)

    def coset_idft_batch(self, mat, shift: int) -> DeviceMatrix:
        return self._dft(mat, shift, True, False)

    def lde_batch(self, mat, added_bits: int) -> DeviceMatrix:
        return self.coset_lde_batch(mat, added_bits, MONTY_ONE)

    def coset_lde_batch(self, mat, added_bits: int, shift: int, bit_reversed: bool = False, out: DeviceMatrix | None = None) -> DeviceMatrix:
        """TwoAdicSubgroupDft::coset_lde_batch.  bit_reversed=True returns what
        `.bit_reverse_rows().to_row_major_matrix()` holds (the layout TwoAdicFriPcs commits to)."""
        m = self._dev(mat)
        if out is not None:
            self.ctx.check(self.ctx.lib.b200zk_coset_lde_batch_into(self.ctx.h, m.h, added_bits, shift, int(bit_reversed), out.h))
            return out
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_coset_lde_batch(self.ctx.h, m.h, added_bits, shift, int(bit_reversed), C.byref(h)))
        return DeviceMatrix(self.ctx, h, True)

    # single-vector conveniences of the trait
    def dft(self, vec):
        import numpy as np
        return self.dft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1)).to_host().reshape(-1)

    def idft(self, vec):
        import numpy as np
        return self.idft_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1)).to_host().reshape(-1)

    def coset_lde(self, vec, added_bits: int, shift: int):
        import numpy as np
        return self.coset_lde_batch(np.asarray(vec, dtype=np.uint32).reshape(-1, 1), added_bits, shift).to_host().reshape(-1)

    # ---- the *_algebra variants of the trait (extension-field inputs).  p3-dft implements them by flattening every extension
    # element into its base coefficients (the transform is F-linear), transforming the `width * D` base columns and
    # reconstituting; an EF4 matrix here is a (rows, width, 4) array or its (rows, 4 * width) flattening, so they are the
    # base-field calls on the flattened view.  FRI's final polynomial (`idft_algebra` of the last folded vector) is the caller.
    @staticmethod
    def _flat(mat):
        import numpy as np
        a = np.ascontiguousarray(mat, dtype=np.uint32)
        if a.ndim == 1:
            raise ValueError("extension-field input needs a trailing axis of 4 coefficients")
        shape = a.shape
        return a.reshape(shape[0], -1), shape

    def dft_algebra_batch(self, mat):
        f, shape = self._flat(mat)
        return self.dft_batch(f).to_host().reshape(shape)

    def idft_algebra_batch(self, mat):
        f, shape = self._flat(mat)
        return self.idft_batch(f).to_host().reshape(shape)

    def coset_dft_algebra_batch(self, mat, shift: int):
        f, shape = self._flat(mat)
        return self.coset_dft_batch(f, shift).to_host().reshape(shape)

    def coset_idft_algebra_batch(self, mat, shift: int):
        f, shape = self._flat(mat)
        return self.coset_idft_batch(f, shift).to_host().reshape(shape)

    def lde_algebra_batch(self, mat, added_bits: int):
        f, shape = self._flat(mat)
        return self.lde_batch(f, added_bits).to_host().reshape((shape[0] << added_bits,) + shape[1:])

    def coset_lde_algebra_batch(self, mat, added_bits: int, shift: int):
        f, shape = self._flat(mat)
        return self.coset_lde_batch(f, added_bits, shift).to_host().reshape((shape[0] << added_bits,) + shape[1:])

    def dft_algebra(self, vec):
        """one vector of EF4 elements, shape (n, 4)"""
        return self.dft_algebra_batch(vec)

    def idft_algebra(self, vec):
        return self.idft_algebra_batch(vec)
