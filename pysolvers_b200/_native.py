"""ctypes binding of libpysolv_b200.so (the C ABI in include/pysolv_b200.h).

The library is the product's only compute path: importing this module on a
machine where it has not been built raises, and no function here has a CPU
fallback.  ``lib()`` loads it lazily so that CPU-only tooling (generators,
partitioner, host-side API tests) can import the package without a GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpysolv_b200.so')

PSB_OK = 0
# status codes of psb_solve_result.status (include/pysolv_b200.h)
CONVERGED, MAXITER, BREAKDOWN_UR, BREAKDOWN_PAP, TRIVIAL, GMRES_FALSE_CONV = range(6)
ORTH_CGS2, ORTH_MGS = 1, 2
SMOOTH_JACOBI, SMOOTH_GS = 0, 1
SPMV_STREAM, SPMV_VECTOR, SPMV_STREAM_LSU, SPMV_MERGE, SPMV_TILE512 = 1, 2, 3, 4, 16


class NativeError(RuntimeError):
    pass


class SolveResult(C.Structure):
    _fields_ = [('status', C.c_int32), ('k', C.c_int32), ('n_hist', C.c_int32),
                ('lucky', C.c_int32), ('norm_r', C.c_double),
                ('norm_b', C.c_double), ('norm_r_rec', C.c_double)]


_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_dbl = C.c_double

# name -> (restype, argtypes); mirrors include/pysolv_b200.h one to one
SIGNATURES = {
    'psb_version': (C.c_int, []),
    'psb_last_error': (C.c_char_p, []),
    'psb_launch_count': (C.c_longlong, []),
    'psb_csr_create': (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    'psb_csr_set_cols16': (C.c_int, [C.c_int]),
    'psb_csr_destroy': (C.c_int, [_vp]),
    'psb_csr_info': (C.c_int, [_vp, C.POINTER(_i64)]),
    'psb_csr_set_kind': (C.c_int, [_vp, C.c_int]),
    'psb_spmv': (C.c_int, [_vp, _vp, _vp, _vp]),
    'psb_spmv_dot': (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'psb_spmv_residual': (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    'psb_spmv_add': (C.c_int, [_vp, _vp, _vp, _vp]),
    'psb_jacobi_sweep': (C.c_int, [_vp, _vp, _dbl, _vp, _vp, _vp, _vp]),
    'psb_dot': (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
    'psb_sa_phase1': (C.c_int, [_i64, _vp, _vp, _vp, _vp, C.POINTER(_i64)]),
    'psb_stencil_nnz': (_i64, [C.c_int, _i64, _i64, _i64]),
    'psb_stencil_fill': (C.c_int, [C.c_int, _i64, _i64, _i64, _dbl, _dbl, _vp, _vp, _vp, _vp]),
    'psb_trsv_create': (C.c_int, [_i64, _vp, _vp, _vp, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    'psb_trsv_destroy': (C.c_int, [_vp]),
    'psb_trsv_info': (C.c_int, [_vp, C.POINTER(_i64)]),
    'psb_trsv_info2': (C.c_int, [_vp, C.POINTER(_i64)]),
    'psb_trsv_set_kernel': (C.c_int, [_vp, C.c_int]),
    'psb_trsv_set_trace': (C.c_int, [_vp, _vp]),
    'psb_trsv_get_levels': (C.c_int, [_vp, _vp, _vp]),
    'psb_trsv_solve': (C.c_int, [_vp, _vp, _vp, _vp]),
    'psb_trsv_error': (C.c_int, [_vp, C.POINTER(_i32)]),
    'psb_ic_create': (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    'psb_ilu_create': (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    'psb_splitlu_create': (C.c_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    'psb_splitlu2_create': (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    'psb_splitlu_set_blockdiag': (C.c_int, [_vp, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    'psb_tri_heights_upper': (C.c_int, [_i64, _vp, _vp, _vp]),
    'psb_tri_levels': (C.c_int, [_i64, _vp, _vp, C.c_int, _vp]),
    'psb_prec_apply': (C.c_int, [_vp, _vp, _vp, _vp]),
    'psb_prec_destroy': (C.c_int, [_vp]),
    'psb_amg_create': (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _dbl, _i32, _i32, _i32, _dbl,
                                 C.POINTER(_vp)]),
    'psb_vec_add': (C.c_int, [_i64, _vp, _vp, _vp]),
    'psb_amg_solve': (C.c_int, [_vp, _vp, _vp, _i32, _dbl, _vp, C.POINTER(SolveResult), _vp]),
    'psb_pcg_workspace_bytes': (_i64, [_i64, C.c_int]),
    'psb_pcg_solve': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _dbl, _i32, _vp,
                                C.POINTER(SolveResult), _vp]),
    'psb_debug_mega_timeline': (C.c_int, [_vp, _i32, _i32]),
    'psb_gmres_workspace_bytes': (_i64, [_i64, _i32]),
    'psb_gmres_solve': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _dbl, _i32, _i32, _vp,
                                  C.POINTER(SolveResult), _vp]),
    'psb_bratu_residual': (C.c_int, [_i64, _vp, _vp, _dbl, _vp, _vp]),
    'psb_bratu_jacobian': (C.c_int, [_i64, _vp, _vp, _vp, _dbl, _vp, _vp]),
    'psb_nccl_unique_id': (C.c_int, [_vp]),
    'psb_comm_create': (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    'psb_comm_destroy': (C.c_int, [_vp]),
    'psb_comm_allreduce_sum': (C.c_int, [_vp, _vp, _i64, _vp]),
    'psb_dist_create': (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.POINTER(_vp)]),
    'psb_dist_destroy': (C.c_int, [_vp]),
    'psb_dist_p2p_alloc': (C.c_int, [_vp, _vp, C.POINTER(_i64)]),
    'psb_dist_p2p_open': (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'psb_dist_spmv': (C.c_int, [_vp, _vp, _vp, _vp]),
    'psb_dist_pcg_workspace_bytes': (_i64, [_i64, _i64]),
    'psb_dist_pcg_solve': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _dbl, _i32, _vp,
                                     C.POINTER(SolveResult), _vp]),
    'psb_dist_amg_create': (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i32, _i32, _i32, _dbl,
                                      C.POINTER(_vp)]),
    'psb_dist_gmres_workspace_bytes': (_i64, [_i64, _i64, _i32]),
    'psb_dist_gmres_solve': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _dbl, _i32, _i32, _vp,
                                       C.POINTER(SolveResult), _vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                'libpysolv_b200.so is not built (%s). Run '
                '`python -m pysolvers_b200.csrc.build`; there is no CPU '
                'fallback for the solve path.' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=''):
    if rc != PSB_OK:
        msg = lib().psb_last_error()
        raise NativeError('%s failed (%d): %s' % (
            what or 'libpysolv_b200 call', rc, msg.decode() if msg else ''))


def launch_count():
    return int(lib().psb_launch_count())
