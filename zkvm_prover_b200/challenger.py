"""Host mirror of p3_challenger::DuplexChallenger<BabyBear, Poseidon2, 16, 8> with the sponge state kept on
the device (so the FRI commit phase can observe/sample without host round trips) + GrindingChallenger."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .device import Context, default_context


class DuplexChallenger:
    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_chal_create(self.ctx.h, C.byref(h)))
        self.h = h

    def observe(self, value):
        v = np.ascontiguousarray(np.atleast_1d(value), dtype=np.uint32).reshape(-1)
        for off in range(0, v.size, 8192):
            part = np.ascontiguousarray(v[off:off + 8192])
            self.ctx.check(self.ctx.lib.b200zk_chal_observe(self.ctx.h, self.h, part.ctypes.data, part.size))

    observe_slice = observe

    def observe_algebra_element(self, ef):
        self.observe(np.asarray(ef, dtype=np.uint32).reshape(-1))

    def sample(self) -> int:
        return int(self.sample_vec(1)[0])

    def sample_vec(self, n: int) -> np.ndarray:
        out = np.empty(n, np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_chal_sample(self.ctx.h, self.h, out.ctypes.data, n))
        return out

    def sample_algebra_element(self) -> np.ndarray:
        return self.sample_vec(4)

    sample_ext_element = sample_algebra_element

    def sample_bits(self, bits: int) -> int:
        out = C.c_uint32()
        self.ctx.check(self.ctx.lib.b200zk_chal_sample_bits(self.ctx.h, self.h, bits, C.byref(out)))
        return int(out.value)

    def grind(self, bits: int) -> int:
        """GrindingChallenger::grind -> canonical witness (the smallest valid one)."""
        out = C.c_uint32()
        self.ctx.check(self.ctx.lib.b200zk_chal_grind(self.ctx.h, self.h, bits, C.byref(out)))
        return int(out.value)

    def state(self) -> np.ndarray:
        out = np.empty(34, np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_chal_state(self.ctx.h, self.h, out.ctypes.data))
        return out

    def set_state(self, state):
        """load the 34 words `state()` returns (e.g. a host DuplexChallenger's sponge state | input buffer | fill | output buffer | fill)"""
        s = np.ascontiguousarray(state, dtype=np.uint32).reshape(34)
        self.ctx.check(self.ctx.lib.b200zk_chal_set_state(self.ctx.h, self.h, s.ctypes.data))
        return self

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.lib.b200zk_chal_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
