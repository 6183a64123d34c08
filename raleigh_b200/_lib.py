"""ctypes binding of libraleigh_b200.so (the C ABI in include/raleigh_b200.h).

There is NO CPU fallback: if the shared library is missing this module raises
at import, and every entry point turns a non-zero return code into
``RuntimeError('cuda error %d: ...')`` -- the reference's convention
(dense_cublas.py:779-781).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libraleigh_b200.so')

RL_F32, RL_F64 = 0, 1

c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_u64 = ctypes.c_uint64

# name -> (restype, argtypes); mirrors include/raleigh_b200.h one to one
PROTOTYPES = {
    'rl_version': (c_int, []),
    'rl_error_string': (ctypes.c_char_p, [c_int]),
    'rl_device_count': (c_int, [ctypes.POINTER(c_int)]),
    'rl_device_info': (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                               ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]),
    'rl_sync_device': (c_int, []),
    'rl_sync_stream': (c_int, [c_vp]),
    'rl_launch_count': (c_i64, []),
    'rl_profile_enable': (None, [c_int]),
    'rl_profile_reset': (None, []),
    'rl_profile_kinds': (c_int, []),
    'rl_profile_name': (ctypes.c_char_p, [c_int]),
    'rl_profile_get': (c_int, [c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl),
                               ctypes.POINTER(c_dbl)]),
    'rl_malloc': (c_int, [ctypes.POINTER(c_vp), c_sz]),
    'rl_free': (c_int, [c_vp]),
    'rl_memset': (c_int, [c_vp, c_int, c_sz, c_vp]),
    'rl_h2d': (c_int, [c_vp, c_vp, c_sz, c_vp]),
    'rl_d2h': (c_int, [c_vp, c_vp, c_sz, c_vp]),
    'rl_h2d_2d': (c_int, [c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp]),
    'rl_d2h_2d': (c_int, [c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp]),
    'rl_copy': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'rl_gather': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, ctypes.POINTER(c_i64), c_i64, c_i64, c_vp]),
    'rl_fill_uniform': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_u64, c_i64, c_i64, c_vp]),
    'rl_axpy': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_dbl, c_vp]),
    'rl_axpy_diag': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_axpy_diag_h': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_scale': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp]),
    'rl_scale_h': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_vp]),
    'rl_dots_ws_bytes': (c_sz, [c_int, c_i64, c_i64]),
    'rl_dots': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    'rl_dots_h': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_dots_t': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_diag_mul': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_gram_ws_bytes': (c_sz, [c_int, c_i64, c_i64, c_i64]),
    'rl_gram': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    'rl_gram_h': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_gram_acc64_ws_bytes': (c_sz, [c_int, c_i64, c_i64, c_i64]),
    'rl_gram_acc64': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    'rl_debug_set_gram_simt': (None, [c_int]),
    'rl_debug_set_update_fma': (None, [c_int]),
    'rl_debug_set_spmm_warps': (None, [c_int]),
    'rl_debug_set_knob': (None, [c_int, c_int]),
    'rl_debug_get_knob': (c_int, [c_int]),
    'rl_update': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_dbl, c_dbl,
                          c_i64, c_vp]),
    'rl_update_h': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_dbl, c_dbl,
                            c_i64, c_vp]),
    'rl_dense_apply': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_int,
                               c_dbl, c_dbl, c_vp]),
    'rl_dense_apply_tc_supported': (c_int, [c_vp, c_i64, c_vp, c_i64]),
    'rl_dense_apply_tc_ws_bytes': (c_sz, [c_i64, c_i64, c_i64, c_int]),
    'rl_split_tf32': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'rl_dense_apply_tc': (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_int,
                                  c_dbl, c_dbl, c_vp, c_sz, c_vp]),
    'rl_minmax_h': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'rl_csr_spmm': (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    'rl_csr_spmm_ex': (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp,
                               c_vp, c_int, c_vp]),
    'rl_spmm_cluster_runs': (c_int, [c_i64, c_vp, c_vp, c_int, c_vp, ctypes.POINTER(c_dbl)]),
    'rl_sell_spmm': (c_int, [c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    'rl_pack_rows': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    'rl_csr_spmm_halo': (c_int, [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp,
                                 c_vp]),
    'rl_sell_spmm_halo': (c_int, [c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64,
                                  c_vp, c_vp]),
    'rl_gram_dev': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    'rl_dots_dev': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_update_dev': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_dbl, c_dbl, c_i64,
                              c_vp]),
    'rl_residual_dev': (c_int, [c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_scale_rsqrt_dev': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    'rl_small_copy': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'rl_small_transpose': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]),
    'rl_small_mirror': (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp]),
    'rl_small_gemm': (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp,
                              c_i64, c_vp]),
    'rl_small_trsm': (c_int, [c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp]),
    'rl_rr_ritz_check': (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'rl_rr_conjugation': (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'rl_rr_piv_chol': (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_dbl, c_vp, c_vp, c_vp]),
    'rl_rr_estimates': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    'rl_rr_select': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    'rl_syevj_cluster_max_n': (c_int, []),
    'rl_syevj_cluster_ws_bytes': (c_sz, [c_i64]),
    'rl_syevj_grid_max_n': (c_int, []),
    'rl_syevj_cluster': (c_int, [c_vp, c_i64, c_i64, c_int, c_dbl, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp, c_vp]),
    'rl_small_set_identity': (c_int, [c_vp, c_i64, c_i64, c_vp]),
    'rl_small_scale_cols': (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
    'rl_psvd_invbound': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    'rl_small_eigh_factor': (c_int, [c_vp, c_i64, c_i64, c_dbl, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp, c_vp]),
    'rl_small_potrf': (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    'rl_small_eigh_ws_bytes': (c_sz, [c_i64]),
    'rl_small_eigh': (c_int, [c_vp, c_i64, c_i64, c_dbl, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp, c_vp]),
    'rl_rr_solve_ws_bytes': (c_sz, [c_i64]),
    'rl_rr_solve': (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64,
                            c_vp, c_vp, c_vp, c_i64, c_dbl, c_vp, c_sz, c_vp, c_vp]),
    'rl_small_to_block': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_i64, c_vp]),
    'rl_block_to_small': (c_int, [c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    'rl_psvd_gershgorin': (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    'rl_psvd_scale': (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    'rl_psvd_coeffs': (c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp]),
    'rl_syevj_ws_bytes': (c_sz, [c_i64]),
    'rl_syevj': (c_int, [c_vp, c_i64, c_vp, c_vp, c_sz, ctypes.POINTER(c_int), c_vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'raleigh_b200: %s is missing -- build it with `python raleigh_b200/build.py` '
            '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    """Raise RuntimeError('cuda error %d') on a non-zero return code."""
    if rc != 0:
        msg = lib.rl_error_string(int(rc))
        raise RuntimeError('cuda error %d: %s' % (rc, msg.decode() if msg else '?'))


def dtype_code(np_type):
    import numpy
    t = numpy.dtype(np_type).type
    if t is numpy.float32:
        return RL_F32
    if t is numpy.float64:
        return RL_F64
    raise ValueError('data type %s not supported' % repr(np_type))
