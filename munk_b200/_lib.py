"""ctypes binding of libmunk_b200.so (the C ABI declared in include/munk_b200.h).

The product path has no CPU fallback: if the library is missing or does not load, importing
this module raises, and every caller fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmunk_b200.so")

c_int, c_long, c_double, c_void_p, c_size_t = (
    ctypes.c_int, ctypes.c_long, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t)
_P = ctypes.POINTER

# name -> (restype, argtypes); mirrors include/munk_b200.h one to one.
SIGNATURES = {
    "munk_last_error_string": (ctypes.c_char_p, []),
    "munk_version": (c_int, []),
    "munk_workspace_create": (c_int, [c_int, _P(c_void_p)]),
    "munk_workspace_destroy": (c_int, [c_void_p]),
    "munk_workspace_stats": (c_int, [c_void_p, _P(c_long), _P(c_double), c_int]),
    "munk_linv_bytes": (c_size_t, [c_int]),
    "munk_assemble_reglap": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_double,
                                     c_void_p, c_long, c_void_p]),
    "munk_assemble_info": (c_int, [c_void_p, _P(c_int), c_void_p]),
    "munk_assemble_reglap_entries": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_double, c_void_p, c_long, c_void_p]),
    "munk_potrf": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_void_p]),
    "munk_potrf_pair": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_void_p, c_int, c_void_p, c_long,
                                c_void_p, c_void_p]),
    "munk_potrf_info": (c_int, [c_void_p, _P(c_int), c_void_p]),
    "munk_potrf_info_async": (c_int, [c_void_p, c_void_p, c_void_p]),
    "munk_trtri": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_void_p]),
    "munk_gram_lower": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    "munk_trsm_left": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_int, c_int, c_void_p,
                               c_long, c_void_p, c_void_p]),
    "munk_trsm_right": (c_int, [c_void_p, c_int, c_void_p, c_long, c_void_p, c_long, c_int, c_void_p,
                                c_void_p]),
    "munk_dgemm_stair": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p,
                                 c_long, c_void_p, c_long, c_double, c_void_p, c_long, c_int,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "munk_transpose": (c_int, [c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    "munk_dgemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_long,
                           c_void_p, c_long, c_double, c_void_p, c_long, c_int, c_void_p]),
    "munk_gather_cols_lower": (c_int, [c_int, c_void_p, c_long, c_int, c_void_p, c_void_p, c_long,
                                       c_void_p]),
    "munk_transpose_lower": (c_int, [c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    "munk_group_mean_rows": (c_int, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_long, c_void_p,
                                     c_long, c_void_p]),
    "munk_gather_pairs": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_void_p]),
    "munk_copy_lower": (c_int, [c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    "munk_download_upper": (c_int, [c_int, c_void_p, c_long, c_void_p, c_long, c_int, c_void_p]),
    "munk_download_block": (c_int, [c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    "munk_download_pageable": (c_int, [c_int, c_long, c_long, c_void_p, c_long, c_void_p, c_long, c_int, c_int,
                                       c_void_p]),
    "munk_host_digest": (c_int, [c_void_p, c_size_t, c_int, c_void_p]),
    "munk_separate_scores": (c_int, [c_int, c_int, c_void_p, c_long] + [c_void_p] * 13 + [c_void_p]),
    "munk_separate_scores_means": (c_int, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_int,
                                           c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                                           c_void_p]),
    "munk_pair_dots": (c_int, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_long, c_void_p, c_long,
                               c_void_p, c_void_p]),
    "munk_pinv_rows_solve": (c_int, [c_void_p, c_int, c_int, c_void_p, c_long, c_int, c_void_p, c_long,
                                     c_void_p, c_long, c_void_p, c_double, _P(c_int), c_void_p, c_void_p]),
    "munk_pinv_rows_work_bytes": (c_size_t, [c_int]),
    "munk_dist_block": (c_int, []),
    "munk_comm_unique_id": (c_int, [c_void_p]),
    "munk_comm_create": (c_int, [c_int, c_int, c_int, c_void_p, c_int, _P(c_void_p)]),
    "munk_comm_destroy": (c_int, [c_void_p]),
    "munk_dist_storage_elems": (c_size_t, [c_int]),
    "munk_dist_create": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, _P(c_void_p)]),
    "munk_dist_destroy": (c_int, [c_void_p]),
    "munk_dist_blocks": (c_int, [c_void_p]),
    "munk_dist_assemble": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "munk_dist_set_rhs": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p]),
    "munk_dist_potrf_begin": (c_int, [c_void_p, c_void_p]),
    "munk_dist_potrf_step": (c_int, [c_void_p, _P(c_int)]),
    "munk_dist_potrf_end": (c_int, [c_void_p]),
    "munk_dist_solve_backward": (c_int, [c_void_p, c_void_p, c_long, c_int, c_void_p]),
    "munk_dist_unpack": (c_int, [c_void_p, c_void_p, c_long, c_void_p]),
    "munk_microbench_dmma": (c_int, [c_int, c_int, _P(c_double)]),
    "munk_microbench_dfma": (c_int, [c_int, c_int, _P(c_double)]),
}


class MunkError(RuntimeError):
    """A non-zero status from the C ABI."""

    def __init__(self, status, where, message):
        super().__init__("%s failed with status %d: %s" % (where, status, message))
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmunk_b200.so is not built (%s). Run `python -m munk_b200.build`; there is no "
            "CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # checked symbol by symbol in tests/test_abi.py
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status, where):
    if status != 0:
        msg = lib.munk_last_error_string()
        raise MunkError(status, where, msg.decode() if msg else "")
    return status
