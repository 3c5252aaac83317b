"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/bb_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  Arrays are numpy uint32 in Montgomery form (p3's in-memory BabyBear).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

P = 0x78000001
HERE = os.path.dirname(os.path.abspath(__file__))
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(native: bool = False) -> str:
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)
    return os.path.join(HERE, "_build", "liboracle_native.so" if native else "liboracle.so")


def _load(native: bool = False):
    name = "liboracle_native.so" if native else "liboracle.so"
    path = os.path.join(HERE, "_build", name)
    src = os.path.join(HERE, "bb_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build(native)
    lib = C.CDLL(path)
    lib.orc_init()
    sig = {
        "orc_to_monty": (C.c_uint32, [C.c_uint32]), "orc_from_monty": (C.c_uint32, [C.c_uint32]),
        "orc_mul": (C.c_uint32, [C.c_uint32, C.c_uint32]), "orc_add": (C.c_uint32, [C.c_uint32, C.c_uint32]),
        "orc_sub": (C.c_uint32, [C.c_uint32, C.c_uint32]), "orc_inv": (C.c_uint32, [C.c_uint32]),
        "orc_pow": (C.c_uint32, [C.c_uint32, C.c_uint64]),
        "orc_two_adic_generator": (C.c_uint32, [C.c_uint32]),
        "orc_ef_mul": (None, [_u32p, _u32p, _u32p]),
        "orc_get_constants": (None, [_u32p, _u32p]),
        "orc_permute_many": (None, [_u32p, C.c_uint64]),
        "orc_hash_rows": (None, [_u32p, C.c_uint64, C.c_uint64, _u32p]),
        "orc_compress_pairs": (None, [_u32p, _u32p, C.c_uint64]),
        "orc_merkle_commit": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32, _u32p, _u32p]),
        "orc_merkle_verify": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32, _u32p, C.c_uint32, C.c_uint64, _u32p]),
        "orc_naive_dft": (None, [_u32p, _u32p, C.c_uint64, C.c_uint64, C.c_int]),
        "orc_dft_batch": (None, [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int]),
        "orc_coset_lde_batch": (None, [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _u32p]),
        "orc_fri_fold": (None, [_u32p, C.c_uint64, _u32p, _u32p]),
        "orc_ef_inv": (None, [_u32p, _u32p]),
        "orc_dot_ext_powers": (None, [_u32p, C.c_uint64, C.c_uint64, _u32p, _u32p]),
        "orc_interpolate_coset_bitrev": (None, [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, _u32p, _u32p]),
        "orc_reduce_openings": (None, [_u32p, C.c_uint64, C.c_uint32, _u32p, _u32p, _u32p, _u32p]),
        "orc_chal_init": (None, [C.c_void_p]), "orc_chal_observe": (None, [C.c_void_p, _u32p, C.c_uint64]),
        "orc_chal_sample": (C.c_uint32, [C.c_void_p]), "orc_chal_sample_ext": (None, [C.c_void_p, _u32p]),
        "orc_chal_sample_bits": (C.c_uint32, [C.c_void_p, C.c_uint32]), "orc_chal_grind": (C.c_uint32, [C.c_void_p, C.c_uint32]),
        "orc_fri_commit_phase": (C.c_uint32, [_u32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, _u32p, _u32p, _u32p]),
        "orc_fill": (None, [_u32p, C.c_uint64, C.c_uint64, C.c_uint64]),
        "orc_checksum": (C.c_uint64, [_u32p, C.c_uint64, C.c_uint64]),
        "orc_fast_available": (C.c_int, []),
        "orc_fast_merkle_commit_single": (C.c_int, [_u32p, C.c_uint64, C.c_uint64, _u32p, _u32p]),
        "orc_fast_coset_lde_batch": (C.c_int, [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _u32p]),
        "orc_fill_strided": (None, [_u32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
        "orc_eval_poly_many": (None, [_u32p, C.c_uint64, _u32p, C.c_uint64, _u32p]),
    }
    for k, (res, args) in sig.items():
        f = getattr(lib, k)
        f.restype, f.argtypes = res, args
    return lib


_LIB = None
_NATIVE = False


def lib(native=None):
    """the loaded oracle library; native=None keeps the current flavour (portable x86-64-v3 build unless use_native(True)
    selected the -march=native one).  r01 bug: the default argument was False, so every plain wrapper call silently switched
    a native selection back to the portable build."""
    global _LIB, _NATIVE
    if native is None:
        native = _NATIVE
    if _LIB is None or native != _NATIVE:
        _LIB, _NATIVE = _load(native), native
    return _LIB


def use_native(flag: bool = True):
    lib(flag)


MONTY_ONE = 0x0FFFFFFE


def to_monty(x):
    x = np.asarray(x, dtype=np.uint64) % P
    return ((x << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def from_monty(m):
    m = np.asarray(m, dtype=np.uint64)
    rinv = pow(1 << 32, -1, P)
    # (m * rinv) fits in 62 bits
    return ((m * np.uint64(rinv)) % np.uint64(P)).astype(np.uint32)


def two_adic_generator(bits: int) -> int:
    return lib().orc_two_adic_generator(bits)


def constants():
    rc = np.zeros(141, np.uint32)
    diag = np.zeros(16, np.uint32)
    lib().orc_get_constants(rc, diag)
    return rc, diag


def permute(states):
    s = np.ascontiguousarray(states, dtype=np.uint32).copy().reshape(-1, 16)
    lib().orc_permute_many(s, s.shape[0])
    return s.reshape(np.shape(states))


def hash_rows(mat):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    out = np.zeros((mat.shape[0], 8), np.uint32)
    lib().orc_hash_rows(mat, mat.shape[0], mat.shape[1], out)
    return out


def compress_pairs(digests):
    d = np.ascontiguousarray(digests, dtype=np.uint32).reshape(-1, 16)
    out = np.zeros((d.shape[0], 8), np.uint32)
    lib().orc_compress_pairs(d, out, d.shape[0])
    return out


def _mat_args(mats):
    mats = [np.ascontiguousarray(m, dtype=np.uint32) for m in mats]
    k = len(mats)
    ptrs = (C.c_void_p * k)(*[m.ctypes.data for m in mats])
    hs = (C.c_uint64 * k)(*[m.shape[0] for m in mats])
    ws = (C.c_uint64 * k)(*[m.shape[1] for m in mats])
    return mats, ptrs, hs, ws, k


def merkle_commit(mats):
    """-> (root[8], digest_layers: list of (len,8) arrays, layer 0 first)."""
    mats, ptrs, hs, ws, k = _mat_args(mats)
    max_h = max(m.shape[0] for m in mats)
    dig = np.zeros((2 * max_h - 1, 8), np.uint32)
    root = np.zeros(8, np.uint32)
    rc = lib().orc_merkle_commit(ptrs, hs, ws, k, dig, root)
    if rc != 0:
        raise ValueError("bad shapes for merkle_commit")
    layers, off, n = [], 0, max_h
    while n >= 1:
        layers.append(dig[off:off + n])
        off += n
        n //= 2
    return root, layers


def merkle_open(mats, layers, index):
    """MerkleTreeMmcs::open_batch: rows at index >> (log max_h - log h) and sibling path."""
    max_h = max(m.shape[0] for m in mats)
    lm = max_h.bit_length() - 1
    rows = [np.asarray(m)[index >> (lm - (m.shape[0].bit_length() - 1))].copy() for m in mats]
    path = np.stack([layers[d][(index >> d) ^ 1] for d in range(lm)]) if lm else np.zeros((0, 8), np.uint32)
    return rows, path


def merkle_verify(rows, heights, path, index, root):
    rows = [np.ascontiguousarray(r, dtype=np.uint32).reshape(1, -1) for r in rows]
    _, ptrs, _, ws, k = _mat_args(rows)
    hs = (C.c_uint64 * k)(*heights)
    path = np.ascontiguousarray(path, dtype=np.uint32).reshape(-1, 8)
    if path.shape[0] == 0:
        path = np.zeros((1, 8), np.uint32)
        depth = 0
    else:
        depth = path.shape[0]
    return bool(lib().orc_merkle_verify(ptrs, hs, ws, k, path, depth, index, np.ascontiguousarray(root, dtype=np.uint32)))


def naive_dft(mat, inverse=False):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    out = np.zeros_like(mat)
    lib().orc_naive_dft(mat, out, mat.shape[0], mat.shape[1], int(inverse))
    return out


def dft_batch(mat, shift=MONTY_ONE, inverse=False, bitrev_out=False):
    a = np.ascontiguousarray(mat, dtype=np.uint32).copy()
    lib().orc_dft_batch(a, a.shape[0], a.shape[1], shift, int(inverse), int(bitrev_out))
    return a


def coset_lde_batch(evals, added_bits, shift, bitrev_out=True):
    evals = np.ascontiguousarray(evals, dtype=np.uint32)
    out = np.zeros((evals.shape[0] << added_bits, evals.shape[1]), np.uint32)
    lib().orc_coset_lde_batch(evals, evals.shape[0], evals.shape[1], added_bits, shift, int(bitrev_out), out)
    return out


def fri_fold(vec, beta):
    vec = np.ascontiguousarray(vec, dtype=np.uint32).reshape(-1, 4)
    out = np.zeros((vec.shape[0] // 2, 4), np.uint32)
    lib().orc_fri_fold(vec, vec.shape[0], np.ascontiguousarray(beta, dtype=np.uint32), out)
    return out


def ef_inv(a):
    out = np.zeros(4, np.uint32)
    lib().orc_ef_inv(np.ascontiguousarray(a, dtype=np.uint32), out)
    return out


def dot_ext_powers(mat, alpha):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    out = np.zeros((mat.shape[0], 4), np.uint32)
    lib().orc_dot_ext_powers(mat, mat.shape[0], mat.shape[1], np.ascontiguousarray(alpha, dtype=np.uint32), out)
    return out


def interpolate_coset_bitrev(evals, shift, point):
    evals = np.ascontiguousarray(evals, dtype=np.uint32)
    out = np.zeros((evals.shape[1], 4), np.uint32)
    lib().orc_interpolate_coset_bitrev(evals, evals.shape[0], evals.shape[1], shift, np.ascontiguousarray(point, dtype=np.uint32), out)
    return out


def reduce_openings(reduced_row, shift, point, reduced_ys, alpha_pow_offset, ro):
    rr = np.ascontiguousarray(reduced_row, dtype=np.uint32).reshape(-1, 4)
    ro = np.ascontiguousarray(ro, dtype=np.uint32).reshape(-1, 4).copy()
    lib().orc_reduce_openings(rr, rr.shape[0], shift, np.ascontiguousarray(point, dtype=np.uint32), np.ascontiguousarray(reduced_ys, dtype=np.uint32),
                              np.ascontiguousarray(alpha_pow_offset, dtype=np.uint32), ro)
    return ro


class Challenger:
    def __init__(self):
        self._buf = C.create_string_buffer(16 * 4 + 8 * 4 + 4 + 8 * 4 + 4)
        lib().orc_chal_init(self._buf)

    def observe(self, v):
        v = np.ascontiguousarray(np.atleast_1d(v), dtype=np.uint32)
        lib().orc_chal_observe(self._buf, v, v.size)

    def sample(self):
        return lib().orc_chal_sample(self._buf)

    def sample_ext(self):
        out = np.zeros(4, np.uint32)
        lib().orc_chal_sample_ext(self._buf, out)
        return out

    def sample_bits(self, bits):
        return lib().orc_chal_sample_bits(self._buf, bits)

    def grind(self, bits):
        return lib().orc_chal_grind(self._buf, bits)

    def state(self):
        return np.frombuffer(self._buf.raw, dtype=np.uint32).copy()


def fri_commit_phase(vec, log_blowup, log_final_poly_len, betas=None, challenger=None):
    """-> (roots (rounds,8), betas (rounds,4), final folded vector (bit-reversed, EF4))."""
    vec = np.ascontiguousarray(vec, dtype=np.uint32).reshape(-1, 4)
    n = vec.shape[0]
    max_rounds = max(n.bit_length(), 1)
    roots = np.zeros((max_rounds, 8), np.uint32)
    bout = np.zeros((max_rounds, 4), np.uint32)
    fin = np.zeros((1 << (log_blowup + log_final_poly_len), 4), np.uint32)
    bptr = None
    if betas is not None:
        betas = np.ascontiguousarray(betas, dtype=np.uint32)
        bptr = betas.ctypes.data
    chal = challenger._buf if challenger is not None else None
    assert bptr is not None or chal is not None
    r = lib().orc_fri_commit_phase(vec, n, log_blowup, log_final_poly_len, bptr, chal, roots, bout, fin)
    return roots[:r], bout[:r], fin


def fill(n, seed, offset=0):
    out = np.zeros(n, np.uint32)
    lib().orc_fill(out, n, seed, offset)
    return out


def fast_available() -> bool:
    """True when the loaded library was compiled with the AVX-512 fast path (`make native` on an AVX-512 host)"""
    return bool(lib().orc_fast_available())


def fast_coset_lde_batch(evals, added_bits, shift):
    """bench.py CPU baseline: coset_lde_batch with bit-reversed rows through the vectorised path (scalar fallback inside)"""
    evals = np.ascontiguousarray(evals, dtype=np.uint32)
    out = np.zeros((evals.shape[0] << added_bits, evals.shape[1]), np.uint32)
    if lib().orc_fast_coset_lde_batch(evals, evals.shape[0], evals.shape[1], added_bits, shift, out) != 0:
        lib().orc_coset_lde_batch(evals, evals.shape[0], evals.shape[1], added_bits, shift, 1, out)
    return out


def fast_merkle_commit_single(mat):
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    dig = np.zeros((2 * mat.shape[0] - 1, 8), np.uint32)
    root = np.zeros(8, np.uint32)
    if lib().orc_fast_merkle_commit_single(mat, mat.shape[0], mat.shape[1], dig, root) != 0:
        raise ValueError("bad shape")
    return root, dig


def fill_column(rows, width, col, seed):
    """column `col` of the rows x width matrix fill(rows * width, seed) would produce, without building the matrix"""
    out = np.zeros(rows, np.uint32)
    lib().orc_fill_strided(out, rows, seed, col, width)
    return out


def eval_poly_many(coef, xs):
    """Horner: the polynomial with Montgomery coefficients `coef` (natural order) at every point of `xs`"""
    coef = np.ascontiguousarray(coef, dtype=np.uint32).reshape(-1)
    xs = np.ascontiguousarray(xs, dtype=np.uint32).reshape(-1)
    out = np.zeros(xs.size, np.uint32)
    lib().orc_eval_poly_many(coef, coef.size, xs, xs.size, out)
    return out


def checksum(v, offset=0):
    v = np.ascontiguousarray(v, dtype=np.uint32).reshape(-1)
    return lib().orc_checksum(v, v.size, offset)
