"""Host mirror of the FRI pieces on the path (v1-era p3-fri): TwoAdicFriPcs::commit (coset LDE of every
trace + one MMCS commit), prover::commit_phase and TwoAdicFriGenericConfig::fold_matrix."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .challenger import DuplexChallenger
from .device import Context, DeviceBuffer, DeviceMatrix, default_context
from .field import GENERATOR_MONTY, P, from_monty, monty_scalar
from .mmcs import DIGEST, MerkleTreeMmcs, ProverData


@dataclass
class FriConfig:
    """p3_fri::FriConfig fields that matter for the commit phase (reference values:
    crates/circuits/chunk-circuit/openvm.toml:1-6 -> log_blowup = 1, num_queries = 100, pow bits 16)."""
    log_blowup: int = 1
    log_final_poly_len: int = 0
    num_queries: int = 100
    proof_of_work_bits: int = 16

    def blowup(self) -> int:
        return 1 << self.log_blowup

    def final_poly_len(self) -> int:
        return 1 << self.log_final_poly_len


def fold_matrix(beta, folded, ctx: Context | None = None, add=None) -> np.ndarray:
    """fold_matrix(beta, (len/2) x 2 EF4 matrix) -> len/2 EF4.  `folded`: (len, 4) uint32, bit-reversed order."""
    ctx = ctx or default_context()
    v = np.ascontiguousarray(folded, dtype=np.uint32).reshape(-1, 4)
    n = v.shape[0]
    d_in, d_out, d_add = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.b200zk_dev_alloc(ctx.h, n * 16, C.byref(d_in)))
    ctx.check(ctx.lib.b200zk_dev_alloc(ctx.h, max(n // 2, 1) * 16, C.byref(d_out)))
    try:
        ctx.check(ctx.lib.b200zk_dev_upload(ctx.h, d_in, v.ctypes.data, n * 16))
        if add is not None:
            a = np.ascontiguousarray(add, dtype=np.uint32).reshape(-1, 4)
            ctx.check(ctx.lib.b200zk_dev_alloc(ctx.h, a.nbytes, C.byref(d_add)))
            ctx.check(ctx.lib.b200zk_dev_upload(ctx.h, d_add, a.ctypes.data, a.nbytes))
        b = np.ascontiguousarray(beta, dtype=np.uint32)
        ctx.check(ctx.lib.b200zk_fri_fold_layer(ctx.h, d_in, n, b.ctypes.data, d_add if add is not None else None, d_out))
        out = np.empty((n // 2, 4), np.uint32)
        ctx.check(ctx.lib.b200zk_dev_download(ctx.h, out.ctypes.data, d_out, out.nbytes))
        return out
    finally:
        for d in (d_in, d_out, d_add):
            if d:
                ctx.lib.b200zk_dev_free(ctx.h, d)


@dataclass
class CommitPhaseResult:
    commits: np.ndarray      # rounds x 8
    data: list               # ProverData per round
    final_poly: np.ndarray   # last folded vector, bit-reversed order (blowup * final_poly_len EF4)
    betas: np.ndarray        # rounds x 4
    final_poly_coeffs: np.ndarray = None  # p3: reverse_slice_index_bits + idft_algebra, truncated to final_poly_len (EF4 coefficients)


def commit_phase(config: FriConfig, inputs, challenger: DuplexChallenger | None, ctx: Context | None = None, betas=None,
                 keep_trees: bool = True) -> CommitPhaseResult:
    """p3_fri::prover::commit_phase.  inputs: list of (len_j, 4) EF4 vectors (bit-reversed order, strictly
    decreasing power-of-two lengths) as host arrays or (device_ptr, len) pairs.  The whole loop (commit,
    observe, sample beta, fold, roll-in) runs on the device; `betas` forces the challenges (tests)."""
    ctx = ctx or default_context()
    lib = ctx.lib
    ptrs, lens, owned = [], [], []
    try:
        for v in inputs:
            if isinstance(v, tuple):
                ptrs.append(v[0])
                lens.append(v[1])
                continue
            a = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1, 4)
            d = C.c_void_p()
            ctx.check(lib.b200zk_dev_alloc(ctx.h, a.nbytes, C.byref(d)))
            owned.append(d)
            ctx.check(lib.b200zk_dev_upload(ctx.h, d, a.ctypes.data, a.nbytes))
            ptrs.append(d.value)
            lens.append(a.shape[0])
        stop = 1 << (config.log_blowup + config.log_final_poly_len)
        max_rounds = max(int(lens[0]).bit_length(), 1)
        roots = np.zeros((max_rounds, DIGEST), np.uint32)
        bout = np.zeros((max_rounds, 4), np.uint32)
        final = np.zeros((stop, 4), np.uint32)
        trees = (C.c_void_p * max_rounds)()
        rounds = C.c_uint32()
        bf = None
        if betas is not None:
            bf = np.ascontiguousarray(betas, dtype=np.uint32)
        ctx.check(lib.b200zk_fri_commit_phase(
            ctx.h, (C.c_void_p * len(ptrs))(*ptrs), (C.c_uint64 * len(lens))(*lens), len(ptrs), config.log_blowup,
            config.log_final_poly_len, challenger.h if challenger is not None else None, bf.ctypes.data if bf is not None else None,
            roots.ctypes.data, bout.ctypes.data, final.ctypes.data, trees if keep_trees else None, C.byref(rounds)))
        r = rounds.value
        data = [ProverData(ctx, C.c_void_p(trees[i]), [DeviceMatrix(ctx, C.c_void_p(lib.b200zk_tree_mat(C.c_void_p(trees[i]), 0)), False)])
                for i in range(r)] if keep_trees else []
        res = CommitPhaseResult(roots[:r].copy(), data, final, bout[:r].copy())
        # final polynomial: un-bit-reverse, iDFT every EF4 coefficient column (idft_algebra), keep final_poly_len coefficients,
        # and let the challenger observe them (p3-fri commit_phase tail)
        lb = stop.bit_length() - 1
        nat = final[[int(format(i, f"0{lb}b")[::-1], 2) if lb else 0 for i in range(stop)]]
        h = C.c_void_p()
        m = ctx.upload(nat)
        ctx.check(lib.b200zk_dft_batch(ctx.h, m.h, 0x0FFFFFFE, 1, 0, C.byref(h)))
        res.final_poly_coeffs = DeviceMatrix(ctx, h, True).to_host()[:config.final_poly_len()]
        if challenger is not None:
            challenger.observe(res.final_poly_coeffs.reshape(-1))
        res._input_buffers = owned  # round-0 leaves alias the first input: keep it alive with the result
        owned = []
        return res
    finally:
        for d in owned:
            lib.b200zk_dev_free(ctx.h, d)


class PendingCommit:
    """a commit whose copies and kernels are enqueued but not waited for (TwoAdicFriPcs.commit_host_async)"""

    def __init__(self, ctx, tree):
        self.ctx, self.tree = ctx, tree

    def result(self):
        root = np.empty(DIGEST, np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_tree_root(self.ctx.h, self.tree, root.ctypes.data))
        n = int(self.ctx.lib.b200zk_tree_num_mats(self.tree))
        ldes = [DeviceMatrix(self.ctx, C.c_void_p(self.ctx.lib.b200zk_tree_mat(self.tree, i)), False) for i in range(n)]
        return root, ProverData(self.ctx, self.tree, ldes)


class TwoAdicFriPcs:
    """The commit half of p3_fri::TwoAdicFriPcs: `commit(evaluations)` = coset LDE of every trace matrix with
    shift GENERATOR / domain_shift (= 31 for the shift-1 trace domains OpenVM uses), rows kept in
    bit-reversed order, then one MerkleTreeMmcs commit over all LDEs.  The LDEs never leave the device."""

    def __init__(self, config: FriConfig | None = None, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.config = config or FriConfig()
        self.mmcs = MerkleTreeMmcs(self.ctx)

    def commit(self, evaluations, domain_shifts=None):
        """evaluations: list of trace matrices (host arrays or DeviceMatrix) on domains shift_i * H_i.
        -> (commitment[8], ProverData holding the bit-reversed LDEs)"""
        if not evaluations:
            raise ValueError("commit needs at least one matrix")
        mats = [m if isinstance(m, DeviceMatrix) else self.ctx.upload(m) for m in evaluations]
        if domain_shifts is None:
            shifts = [GENERATOR_MONTY] * len(mats)
        else:  # shift = GENERATOR / domain.shift  (canonical ints in, Montgomery out)
            shifts = [monty_scalar(31 * pow(int(s), -1, P) % P) for s in domain_shifts]
        arr = (C.c_void_p * len(mats))(*[m.h for m in mats])
        sh = (C.c_uint32 * len(mats))(*shifts)
        root = np.empty(DIGEST, np.uint32)
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_lde_commit(self.ctx.h, arr, len(mats), self.config.log_blowup, sh, root.ctypes.data, C.byref(t)))
        n = int(self.ctx.lib.b200zk_tree_num_mats(t))
        ldes = [DeviceMatrix(self.ctx, C.c_void_p(self.ctx.lib.b200zk_tree_mat(t, i)), False) for i in range(n)]
        return root, ProverData(self.ctx, t, ldes)

    def commit_host(self, trace, strip_cols: int = 0, host_ptr: int | None = None, shape=None):
        """commit of ONE host-resident trace with the PCIe transfer overlapped with the arithmetic (column-strip
        pipeline, b200zk_lde_commit_host).  `trace`: C-contiguous uint32 array (pinned memory for full speed), or pass
        `host_ptr` + `shape` for a raw pinned buffer (e.g. a torch pinned tensor's data_ptr)."""
        if host_ptr is None:
            a = np.ascontiguousarray(trace, dtype=np.uint32)
            host_ptr, shape = a.ctypes.data, a.shape
        root = np.empty(DIGEST, np.uint32)
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_lde_commit_host(self.ctx.h, host_ptr, shape[0], shape[1], self.config.log_blowup, GENERATOR_MONTY, strip_cols,
                                                           root.ctypes.data, C.byref(t)))
        n = int(self.ctx.lib.b200zk_tree_num_mats(t))
        ldes = [DeviceMatrix(self.ctx, C.c_void_p(self.ctx.lib.b200zk_tree_mat(t, i)), False) for i in range(n)]
        return root, ProverData(self.ctx, t, ldes)

    def commit_host_async(self, host_ptr: int, shape, strip_cols: int = 0) -> "PendingCommit":
        """b200zk_lde_commit_host_async: enqueue the whole strip pipeline and return; `result()` of the handle waits and
        gives (root, ProverData).  The host buffer must stay untouched until then.  Two may be in flight per context: a
        prover walking the segments of a chunk proof issues segment i + 1 before collecting segment i."""
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.b200zk_lde_commit_host_async(self.ctx.h, host_ptr, shape[0], shape[1], self.config.log_blowup, GENERATOR_MONTY, strip_cols, C.byref(t)))
        return PendingCommit(self.ctx, t)

    # ---- open phase: the data-parallel body of TwoAdicFriPcs::open, LDEs stay on the device
    def inv_denominators(self, log_height: int, point) -> DeviceBuffer:
        """1 / (z - x) for every x of the (bit-reversed) LDE domain GENERATOR * <w_(2^log_height)>"""
        buf = DeviceBuffer(self.ctx, 16 << log_height)
        z = np.ascontiguousarray(point, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_open_denominators(self.ctx.h, log_height, GENERATOR_MONTY, z.ctypes.data, buf.ptr))
        return buf

    def dot_ext_powers(self, lde: DeviceMatrix, alpha) -> DeviceBuffer:
        """Matrix::dot_ext_powers(alpha): one EF4 per LDE row"""
        buf = DeviceBuffer(self.ctx, 16 * lde.rows)
        a = np.ascontiguousarray(alpha, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_mat_dot_ext_powers(self.ctx.h, lde.h, a.ctypes.data, buf.ptr))
        return buf

    def interpolate_coset(self, lde: DeviceMatrix, point, inv_den: DeviceBuffer) -> np.ndarray:
        """opened values p_c(z) of every column (width x 4), from the low coset of the committed LDE"""
        ys = np.empty((lde.width, 4), np.uint32)
        z = np.ascontiguousarray(point, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_interpolate_coset(self.ctx.h, lde.h, self.config.log_blowup, GENERATOR_MONTY, z.ctypes.data, inv_den.ptr,
                                                            ys.ctypes.data))
        return ys

    def reduce_openings(self, reduced_row: DeviceBuffer, m: int, inv_den: DeviceBuffer, reduced_ys, alpha_pow_offset, ro: DeviceBuffer):
        """ro[i] += alpha^offset * (reduced_ys - reduced_row[i]) / (z - x_i)"""
        rys = np.ascontiguousarray(reduced_ys, dtype=np.uint32)
        apo = np.ascontiguousarray(alpha_pow_offset, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.b200zk_reduce_openings(self.ctx.h, reduced_row.ptr, m, inv_den.ptr, rys.ctypes.data, apo.ctypes.data, ro.ptr))

    def open(self, rounds, challenger: DuplexChallenger):
        """The prover side of p3_fri::TwoAdicFriPcs::open, with every data-parallel step on the device.

        rounds: list of (ProverData, points) where points[i] is the list of EF4 opening points of matrix i of that
        commitment (e.g. [zeta, zeta * g_trace]).  Returns (opened_values, proof):
          opened_values[round][matrix][point] = (width, 4) array of p_c(z)
          proof = {commit_phase_commits, final_poly, pow_witness, query_indices, input_openings, commit_phase_openings}
        Composition (alpha sampled first, per-height reduced openings with running alpha offsets, inputs rolled into
        the FRI folding at their height, PoW, then queries) follows p3-fri as of the reference's fixture era; the order
        of transcript operations is restated from memory of that crate and is NOT pinned by a reference vector
         -- the arithmetic of every step is."""
        alpha = challenger.sample_algebra_element()
        reduced, num_reduced, keep = {}, {}, []
        # alpha^0 .. alpha^n on the device, n = the largest running offset any height class reaches
        per_height, total_cols = {}, 0
        for pd, points in rounds:
            for lde, pts in zip(pd.mats, points):
                per_height[lde.rows] = per_height.get(lde.rows, 0) + lde.width * len(pts)
                total_cols += lde.width * len(pts)
        n_pows = max(per_height.values()) + 1
        a4 = np.ascontiguousarray(alpha, dtype=np.uint32)
        alpha_pows = DeviceBuffer(self.ctx, 16 * n_pows)
        self.ctx.check(self.ctx.lib.b200zk_ext_powers(self.ctx.h, a4.ctypes.data, n_pows, alpha_pows.ptr))
        ys_all = DeviceBuffer(self.ctx, 16 * total_cols)   # every opened value, downloaded once at the end
        inv_cache, slots, off = {}, [], 0
        for pd, points in rounds:
            per_round = []
            for lde, pts in zip(pd.mats, points):
                lh = lde.rows.bit_length() - 1
                if lh not in reduced:
                    reduced[lh] = DeviceBuffer(self.ctx, 16 * lde.rows).zero()
                    num_reduced[lh] = 0
                rr = self.dot_ext_powers(lde, alpha)
                per_mat = []
                for zpt in pts:
                    z4 = np.ascontiguousarray(zpt, dtype=np.uint32)
                    key = (lh, z4.tobytes())
                    if key not in inv_cache:               # matrices of one height share 1 / (z - x) for a common point
                        inv_cache[key] = self.inv_denominators(lh, zpt)
                    # opened values -> reduced opening -> accumulate into the height's FRI input, all on the device
                    self.ctx.check(self.ctx.lib.b200zk_open_reduce(self.ctx.h, lde.h, self.config.log_blowup, GENERATOR_MONTY, z4.ctypes.data,
                                                                   inv_cache[key].ptr, rr.ptr, alpha_pows.ptr, num_reduced[lh], reduced[lh].ptr,
                                                                   ys_all.ptr + 16 * off))
                    num_reduced[lh] += lde.width
                    per_mat.append((off, lde.width))
                    off += lde.width
                keep.append(rr)
                per_round.append(per_mat)
            slots.append(per_round)
        heights = sorted(reduced, reverse=True)
        inputs = [(reduced[lh].ptr, 1 << lh) for lh in heights]
        res = commit_phase(self.config, inputs, challenger, self.ctx)
        pow_witness = challenger.grind(self.config.proof_of_work_bits)
        log_max = heights[0]
        # sample_bits = canonical value of one sampled element, masked: all query indices in one device round trip
        indices = [int(v) & ((1 << log_max) - 1) for v in from_monty(challenger.sample_vec(self.config.num_queries))]
        input_openings = []
        for pd, _ in rounds:
            lmh = max(m.rows for m in pd.mats).bit_length() - 1
            input_openings.append(self.mmcs.open_batch_many([i >> (log_max - lmh) for i in indices], pd))
        # commit-phase openings of every round in one call (one download instead of a synchronisation per round)
        cp_openings = []
        if res.data:
            nq, depths = len(indices), [t.depth for t in res.data]
            idx = np.ascontiguousarray(indices, dtype=np.uint64)
            pairs = np.empty((len(res.data), nq, 8), np.uint32)
            paths = np.empty(8 * nq * sum(depths), np.uint32)
            arr = (C.c_void_p * len(res.data))(*[t.h for t in res.data])
            self.ctx.check(self.ctx.lib.b200zk_fri_open_queries(self.ctx.h, arr, len(res.data), idx.ctypes.data, nq, pairs.ctypes.data, paths.ctypes.data))
            off = 0
            for r, d in enumerate(depths):
                pr = paths[off:off + 8 * nq * d].reshape(nq, d, 8)
                off += 8 * nq * d
                cp_openings.append([(pairs[r, q].reshape(2, 4), pr[q]) for q in range(nq)])
        proof = {"alpha": alpha, "commit_phase_commits": res.commits, "betas": res.betas, "final_poly": res.final_poly_coeffs,
                 "pow_witness": pow_witness, "query_indices": indices, "input_openings": input_openings, "commit_phase_openings": cp_openings,
                 "log_max_height": log_max}
        ys_host = ys_all.to_host((total_cols, 4))
        opened = [[[ys_host[o:o + w] for o, w in per_mat] for per_mat in per_round] for per_round in slots]
        res._keep = (reduced, keep, inv_cache, alpha_pows, ys_all)
        proof["_commit_phase"] = res
        return opened, proof

    def get_evaluations_on_domain(self, prover_data: ProverData, idx: int) -> DeviceMatrix:
        return prover_data.mats[idx]
