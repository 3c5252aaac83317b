"""Host mirror of p3_commit::Mmcs as implemented by p3_merkle_tree::MerkleTreeMmcs<.., 8> and
p3_commit::ExtensionMmcs (v1-era crates behind the reference's fixtures).  commit / open_batch /
get_matrices / verify_batch keep the trait's argument meaning; ProverData lives on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .device import Context, DeviceMatrix, default_context

DIGEST = 8


class ProverData:
    """MerkleTree<..>: leaves (matrices) + digest layers, device resident."""

    def __init__(self, ctx: Context, handle, mats):
        self.ctx, self.h, self.mats = ctx, handle, mats

    @property
    def depth(self) -> int:
        return int(self.ctx.lib.b200zk_tree_depth(self.h))

    def root(self) -> np.ndarray:
        out = np.empty(DIGEST, np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_tree_root(self.ctx.h, self.h, out.ctypes.data))
        return out

    def layer(self, i: int) -> np.ndarray:
        n = (1 << self.depth) >> i
        out = np.empty((n, DIGEST), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_tree_download_layer(self.ctx.h, self.h, i, out.ctypes.data))
        return out

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.lib.b200zk_tree_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MerkleTreeMmcs:
    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    def commit(self, inputs):
        """Mmcs::commit(Vec<M>) -> (Commitment, ProverData).  Takes ownership of the matrices."""
        if not inputs:
            raise ValueError("commit needs at least one matrix")
        mats = [m if isinstance(m, DeviceMatrix) else self.ctx.upload(m) for m in inputs]
        arr = (C.c_void_p * len(mats))(*[m.h for m in mats])
        root = np.empty(DIGEST, np.uint32)
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_merkle_commit(self.ctx.h, arr, len(mats), 0, root.ctypes.data, C.byref(t)))
        return root, ProverData(self.ctx, t, mats)  # python keeps the matrix handles alive with the ProverData

    def commit_matrix(self, m):
        return self.commit([m])

    def open_batch(self, index: int, prover_data: ProverData):
        """-> (opened_values: list of rows, opening_proof: (depth, 8) siblings bottom-up)"""
        total = int(self.ctx.lib.b200zk_tree_total_width(prover_data.h))
        rows = np.empty(total, np.uint32)
        path = np.empty((prover_data.depth, DIGEST), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_merkle_open(self.ctx.h, prover_data.h, index, rows.ctypes.data, path.ctypes.data))
        out, off = [], 0
        for m in prover_data.mats:
            out.append(rows[off:off + m.width].copy())
            off += m.width
        return out, path

    def open_batch_many(self, indices, prover_data: ProverData):
        """the query phase's loop over `open_batch` in one launch: -> list of (opened_values, opening_proof)"""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        total = int(self.ctx.lib.b200zk_tree_total_width(prover_data.h))
        depth = prover_data.depth
        rows = np.empty((idx.size, total), np.uint32)
        paths = np.empty((idx.size, depth, DIGEST), np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_merkle_open_many(self.ctx.h, prover_data.h, idx.ctypes.data, idx.size, rows.ctypes.data, paths.ctypes.data))
        bounds = np.cumsum([0] + [m.width for m in prover_data.mats])
        return [([rows[q, a:b] for a, b in zip(bounds[:-1], bounds[1:])], paths[q]) for q in range(idx.size)]   # views into one buffer

    def get_matrices(self, prover_data: ProverData):
        return list(prover_data.mats)

    def get_max_height(self, prover_data: ProverData) -> int:
        return max(m.rows for m in prover_data.mats)

    def verify_batch(self, commit, dimensions, index: int, opened_values, opening_proof) -> None:
        """dimensions: list of (width, height).  Raises ValueError on mismatch (Err(RootMismatch / WrongBatchSize...))."""
        if len(dimensions) != len(opened_values):
            raise ValueError("WrongBatchSize")
        widths = np.array([d[0] for d in dimensions], np.uint32)
        heights = np.array([d[1] for d in dimensions], np.uint64)
        for w, r in zip(widths, opened_values):
            if len(r) != w:
                raise ValueError("WrongWidth")
        rows = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint32).reshape(-1) for r in opened_values]))
        path = np.ascontiguousarray(opening_proof, dtype=np.uint32).reshape(-1, DIGEST)
        if path.shape[0] != int(max(heights)).bit_length() - 1:
            raise ValueError("WrongHeight")
        root = np.ascontiguousarray(commit, dtype=np.uint32)
        ok = C.c_int(0)
        pp = path.ctypes.data if path.shape[0] else None
        self.ctx.check(self.ctx.lib.b200zk_merkle_verify(self.ctx.h, rows.ctypes.data, heights.ctypes.data, widths.ctypes.data, len(widths),
                                                         pp, path.shape[0], index, root.ctypes.data, C.byref(ok)))
        if not ok.value:
            raise ValueError("RootMismatch")


class ExtensionMmcs:
    """p3_commit::ExtensionMmcs<BabyBear, EF4, MerkleTreeMmcs>: EF4 matrices are committed as base matrices
    4x wider (each extension element flattened to its 4 coefficients)."""

    def __init__(self, inner: MerkleTreeMmcs | None = None):
        self.inner = inner or MerkleTreeMmcs()

    @staticmethod
    def _flatten(m):
        a = np.ascontiguousarray(m, dtype=np.uint32)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("expected rows x width x 4 (EF4 coefficients)")
        return a.reshape(a.shape[0], a.shape[1] * 4)

    def commit(self, inputs):
        return self.inner.commit([m if isinstance(m, DeviceMatrix) else self._flatten(m) for m in inputs])

    def commit_matrix(self, m):
        return self.commit([m])

    def open_batch(self, index, prover_data):
        rows, path = self.inner.open_batch(index, prover_data)
        return [r.reshape(-1, 4) for r in rows], path

    def verify_batch(self, commit, dimensions, index, opened_values, opening_proof):
        self.inner.verify_batch(commit, [(w * 4, h) for w, h in dimensions], index, [np.asarray(r).reshape(-1) for r in opened_values], opening_proof)
