"""Context and device-matrix handles over the C ABI (b200zk_ctx / b200zk_mat)."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import B200zkError


class Context:
    """One per (thread, GPU): stream + twiddle caches.  Mirrors the role of the engine object the
    reference selects in crates/prover/src/prover/mod.rs:27-39."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.b200zk_ctx_create(device, C.byref(h))
        if rc != 0:
            raise B200zkError(rc, f"cannot create a context on CUDA device {device} (no CPU fallback exists)")
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise B200zkError(rc, self.lib.b200zk_last_error(self.h).decode())

    def sync(self):
        self.check(self.lib.b200zk_ctx_sync(self.h))

    def trim(self):
        """return pooled device memory to the driver"""
        self.check(self.lib.b200zk_ctx_trim(self.h))

    @property
    def stream(self) -> int:
        return self.lib.b200zk_ctx_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(self.lib.b200zk_kernel_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200zk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- matrices
    def upload(self, arr) -> "DeviceMatrix":
        a = np.ascontiguousarray(arr, dtype=np.uint32)
        if a.ndim != 2:
            raise ValueError("expected a rows x width matrix")
        h = C.c_void_p()
        self.check(self.lib.b200zk_mat_upload(self.h, a.ctypes.data, a.shape[0], a.shape[1], C.byref(h)))
        return DeviceMatrix(self, h, True)

    def alloc(self, rows: int, width: int) -> "DeviceMatrix":
        h = C.c_void_p()
        self.check(self.lib.b200zk_mat_alloc(self.h, rows, width, C.byref(h)))
        return DeviceMatrix(self, h, True)

    def wrap(self, dev_ptr: int, rows: int, width: int, keepalive=None) -> "DeviceMatrix":
        """borrow device memory owned by someone else (e.g. torch_tensor.data_ptr())"""
        h = C.c_void_p()
        self.check(self.lib.b200zk_mat_wrap(self.h, dev_ptr, rows, width, C.byref(h)))
        m = DeviceMatrix(self, h, True)
        m._keepalive = keepalive
        return m


_default = threading.local()


def default_context(device: int = 0) -> Context:
    ctxs = getattr(_default, "ctxs", None)
    if ctxs is None:
        ctxs = _default.ctxs = {}
    if device not in ctxs:
        ctxs[device] = Context(device)
    return ctxs[device]


class DeviceBuffer:
    """raw device allocation from the library's pool (EF4 vectors of the open phase, FRI inputs, ...)"""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx, self.nbytes = ctx, nbytes
        p = C.c_void_p()
        ctx.check(ctx.lib.b200zk_dev_alloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, ctx: Context, arr) -> "DeviceBuffer":
        a = np.ascontiguousarray(arr)
        b = cls(ctx, a.nbytes)
        ctx.check(ctx.lib.b200zk_dev_upload(ctx.h, b.ptr, a.ctypes.data, a.nbytes))
        return b

    def zero(self):
        self.ctx.check(self.ctx.lib.b200zk_dev_zero(self.ctx.h, self.ptr, self.nbytes))
        return self

    def to_host(self, shape, dtype=np.uint32) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx.check(self.ctx.lib.b200zk_dev_download(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr and self.ctx.h:
            self.ctx.lib.b200zk_dev_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceMatrix:
    """RowMajorMatrix<BabyBear> resident in HBM."""

    def __init__(self, ctx: Context, handle, owns_handle: bool):
        self.ctx, self.h, self._owns = ctx, handle, owns_handle
        self._keepalive = None
        self._rows = self._width = None   # a matrix never changes shape: ask the library once

    @property
    def rows(self) -> int:
        if self._rows is None:
            self._rows = int(self.ctx.lib.b200zk_mat_rows(self.h))
        return self._rows

    height = rows

    @property
    def width(self) -> int:
        if self._width is None:
            self._width = int(self.ctx.lib.b200zk_mat_width(self.h))
        return self._width

    @property
    def device_ptr(self) -> int:
        return self.ctx.lib.b200zk_mat_device_ptr(self.h) or 0

    def to_host(self) -> np.ndarray:
        out = np.empty((self.rows, self.width), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_mat_download(self.ctx.h, self.h, out.ctypes.data))
        return out

    def rows_to_host(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.width), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_mat_download_rows(self.ctx.h, self.h, row0, nrows, out.ctypes.data))
        return out

    def fill(self, seed: int):
        self.ctx.check(self.ctx.lib.b200zk_mat_fill(self.ctx.h, self.h, seed))
        return self

    def checksum(self) -> int:
        out = C.c_uint64()
        self.ctx.check(self.ctx.lib.b200zk_mat_checksum(self.ctx.h, self.h, C.byref(out)))
        return int(out.value)

    def release(self):
        """give up the handle without freeing (ownership moved into a tree)"""
        self._owns = False

    def free(self):
        if self._owns and self.h and self.ctx.h:
            self.ctx.lib.b200zk_mat_free(self.ctx.h, self.h)
        self.h = None
        self._owns = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
