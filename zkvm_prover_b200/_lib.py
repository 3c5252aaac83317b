"""ctypes binding of libb200zk.so (C ABI in include/b200zk.h).

There is no CPU path: if the shared library is missing, or no CUDA device is present, everything here
raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ZK_LIB_PATH", os.path.join(HERE, "libb200zk.so"))  # override only for A/B kernel experiments
HEADER = os.path.join(os.path.dirname(HERE), "include", "b200zk.h")

OK, ERR_CUDA, ERR_OOM, ERR_SHAPE, ERR_ARG = 0, -1, -2, -3, -4
_ERR_NAMES = {ERR_CUDA: "B200ZK_ERR_CUDA", ERR_OOM: "B200ZK_ERR_OOM", ERR_SHAPE: "B200ZK_ERR_SHAPE", ERR_ARG: "B200ZK_ERR_ARG"}


class B200zkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def declared_symbols(header: str = HEADER):
    """Every function the public header declares (used by the export test)."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", text)))


_p = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_u64 = C.c_uint64
_u32 = C.c_uint32
_int = C.c_int

_SIGS = {
    "b200zk_version": (C.c_char_p, []),
    "b200zk_ctx_create": (_int, [_int, C.POINTER(_p)]),
    "b200zk_ctx_destroy": (None, [_p]),
    "b200zk_last_error": (C.c_char_p, [_p]),
    "b200zk_ctx_sync": (_int, [_p]),
    "b200zk_ctx_stream": (_p, [_p]),
    "b200zk_ctx_trim": (_int, [_p]),
    "b200zk_kernel_launches": (_u64, [_p]),
    "b200zk_mat_alloc": (_int, [_p, _u64, _u32, C.POINTER(_p)]),
    "b200zk_mat_upload": (_int, [_p, _p, _u64, _u32, C.POINTER(_p)]),
    "b200zk_mat_upload_into": (_int, [_p, _p, _p]),
    "b200zk_mat_wrap": (_int, [_p, _p, _u64, _u32, C.POINTER(_p)]),
    "b200zk_mat_download": (_int, [_p, _p, _p]),
    "b200zk_mat_download_rows": (_int, [_p, _p, _u64, _u64, _p]),
    "b200zk_mat_rows": (_u64, [_p]),
    "b200zk_mat_width": (_u32, [_p]),
    "b200zk_mat_device_ptr": (_p, [_p]),
    "b200zk_mat_free": (None, [_p, _p]),
    "b200zk_mat_fill": (_int, [_p, _p, _u64]),
    "b200zk_mat_checksum": (_int, [_p, _p, C.POINTER(_u64)]),
    "b200zk_coset_lde_batch": (_int, [_p, _p, _u32, _u32, _int, C.POINTER(_p)]),
    "b200zk_coset_lde_batch_into": (_int, [_p, _p, _u32, _u32, _int, _p]),
    "b200zk_dft_batch": (_int, [_p, _p, _u32, _int, _int, C.POINTER(_p)]),
    "b200zk_poseidon2_permute": (_int, [_p, _p, _u64]),
    "b200zk_poseidon2_permute_dev": (_int, [_p, _p, _u64]),
    "b200zk_poseidon2_permute_plain_dev": (_int, [_p, _p, _u64]),
    "b200zk_hash_rows": (_int, [_p, _p, _p]),
    "b200zk_hash_rows_dev": (_int, [_p, _p, _p]),
    "b200zk_compress_pairs": (_int, [_p, _p, _p, _u64]),
    "b200zk_compress_pairs_dev": (_int, [_p, _p, _p, _u64]),
    "b200zk_merkle_commit": (_int, [_p, C.POINTER(_p), _u32, _int, _p, C.POINTER(_p)]),
    "b200zk_lde_commit": (_int, [_p, C.POINTER(_p), _u32, _u32, _p, _p, C.POINTER(_p)]),
    "b200zk_lde_commit_host": (_int, [_p, _p, _u64, _u32, _u32, _u32, _u32, _p, C.POINTER(_p)]),
    "b200zk_lde_commit_host_async": (_int, [_p, _p, _u64, _u32, _u32, _u32, _u32, C.POINTER(_p)]),
    "b200zk_host_alloc": (_int, [_u64, _int, C.POINTER(_p)]),
    "b200zk_host_free": (None, [_p]),
    "b200zk_merkle_open": (_int, [_p, _p, _u64, _p, _p]),
    "b200zk_merkle_open_many": (_int, [_p, _p, _p, _u32, _p, _p]),
    "b200zk_tree_depth": (_u32, [_p]),
    "b200zk_tree_num_mats": (_u32, [_p]),
    "b200zk_tree_total_width": (_u64, [_p]),
    "b200zk_tree_mat": (_p, [_p, _u32]),
    "b200zk_tree_root": (_int, [_p, _p, _p]),
    "b200zk_tree_download_layer": (_int, [_p, _p, _u32, _p]),
    "b200zk_tree_free": (None, [_p, _p]),
    "b200zk_merkle_verify": (_int, [_p, _p, _p, _p, _u32, _p, _u32, _u64, _p, C.POINTER(_int)]),
    "b200zk_chal_create": (_int, [_p, C.POINTER(_p)]),
    "b200zk_chal_free": (None, [_p, _p]),
    "b200zk_chal_observe": (_int, [_p, _p, _p, _u32]),
    "b200zk_chal_sample": (_int, [_p, _p, _p, _u32]),
    "b200zk_chal_sample_bits": (_int, [_p, _p, _u32, _p]),
    "b200zk_chal_grind": (_int, [_p, _p, _u32, _p]),
    "b200zk_chal_state": (_int, [_p, _p, _p]),
    "b200zk_chal_set_state": (_int, [_p, _p, _p]),
    "b200zk_fri_commit_layer": (_int, [_p, _p, _u64, _p, C.POINTER(_p)]),
    "b200zk_fri_fold_layer": (_int, [_p, _p, _u64, _p, _p, _p]),
    "b200zk_fri_commit_phase": (_int, [_p, C.POINTER(_p), C.POINTER(_u64), _u32, _u32, _u32, _p, _p, _p, _p, _p, C.POINTER(_p), C.POINTER(_u32)]),
    "b200zk_open_denominators": (_int, [_p, _u32, _u32, _p, _p]),
    "b200zk_mat_dot_ext_powers": (_int, [_p, _p, _p, _p]),
    "b200zk_interpolate_coset": (_int, [_p, _p, _u32, _u32, _p, _p, _p]),
    "b200zk_reduce_openings": (_int, [_p, _p, _u64, _p, _p, _p, _p]),
    "b200zk_dev_alloc": (_int, [_p, _u64, C.POINTER(_p)]),
    "b200zk_dev_free": (None, [_p, _p]),
    "b200zk_dev_upload": (_int, [_p, _p, _p, _u64]),
    "b200zk_dev_zero": (_int, [_p, _p, _u64]),
    "b200zk_fri_open_queries": (_int, [_p, _p, _u32, _p, _u32, _p, _p]),
    "b200zk_peer_alloc": (_int, [_p, _u64, _p, _p]),
    "b200zk_peer_open": (_int, [_p, _p, _p]),
    "b200zk_peer_close": (_int, [_p, _p]),
    "b200zk_peer_free": (_int, [_p, _p]),
    "b200zk_coset_lde_scatter": (_int, [_p, _p, _u32, _u32, _u32, _u32, _p]),
    "b200zk_coset_lde_scatter_rows": (_int, [_p, _p, _u32, _u32, _u32, _u32, _p]),
    "b200zk_ext_powers": (_int, [_p, _p, _u32, _p]),
    "b200zk_open_reduce": (_int, [_p, _p, _u32, _u32, _p, _p, _p, _p, _u32, _p, _p]),
    "b200zk_dev_download": (_int, [_p, _p, _p, _u64]),
}

_LIB = None


def load():
    """dlopen libb200zk.so and type every entry point.  Raises if the library was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        f = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        f.restype, f.argtypes = res, args
    _LIB = lib
    return lib
