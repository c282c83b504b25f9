"""ctypes binding of libpsum.so (include/psum.h).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PSUM_LIB_PATH') or os.path.join(HERE, 'libpsum.so')  # override: kernel-variant experiments

PSUM_OK = 0
ERR_INVALID, ERR_CUDA, ERR_EMPTY, ERR_NOT_HERMITIAN, ERR_NOT_DIAGONAL, ERR_NCCL, ERR_NOMEM, \
    ERR_NOT_CONVERGED = -1, -2, -3, -4, -5, -6, -7, -8

_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

SIGNATURES = {
    'psum_last_error': (C.c_char_p, []),
    'psum_version': (C.c_int, []),
    'psum_create': (C.c_int, [C.c_int, C.c_int64, _u64p, _u64p, C.POINTER(C.c_uint8), _dp,
                              C.POINTER(_vp)]),
    'psum_destroy': (C.c_int, [_vp]),
    'psum_set_option': (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    'psum_info': (C.c_int, [_vp, _i64p, C.c_int]),
    'psum_pass_info': (C.c_int, [_vp, C.c_int, C.c_int, _u64p, _i64p, _i64p, _i64p]),
    'psum_plan_export': (C.c_int, [_vp, C.c_int, C.c_int, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'psum_plan_floors': (C.c_int, [_vp, C.c_int, _dp]),
    'psum_exchange_stats': (C.c_int, [_vp, _dp]),
    'psum_apply': (C.c_int, [_vp, _vp, _vp, _dp, _vp]),
    'psum_expect': (C.c_int, [_vp, _vp, _dp, _vp]),
    'psum_expect_host': (C.c_int, [_vp, _vp, _vp, _dp, _vp]),
    'psum_bra_apply': (C.c_int, [_vp, _vp, _vp, _dp, _vp]),
    'psum_expect_batch': (C.c_int, [_vp, C.c_int, _vp, _dp, _vp]),
    'psum_to_dense': (C.c_int, [_vp, _vp, _vp]),
    'psum_pauli_decompose': (C.c_int, [_vp, C.c_int, _vp, _vp]),
    'psum_expect_multi': (C.c_int, [C.POINTER(_vp), C.c_int, _vp, _dp, _vp]),
    'psum_diag': (C.c_int, [_vp, _vp, _vp]),
    'psum_diag_kmin': (C.c_int, [_vp, C.c_int, _dp, _u64p, _vp]),
    'psum_lanczos': (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_int64, _dp, _vp,
                               _dp, C.POINTER(C.c_int), _vp]),
    'psum_arnoldi': (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_int64, _dp, _vp,
                               _dp, C.POINTER(C.c_int), _vp]),
    'psum_evolve': (C.c_int, [_vp, _vp, _vp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int),
                              C.POINTER(C.c_int), _vp]),
    'psum_state_basis': (C.c_int, [_vp, C.c_int, C.c_uint64, _vp]),
    'psum_gate_1q': (C.c_int, [_vp, C.c_int, C.c_int, _dp, _vp]),
    'psum_gate_2q': (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'psum_gate_block': (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int32), C.c_int, _vp, C.c_int, _vp]),
    'psum_fill_state': (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]),
    'psum_vec_dot': (C.c_int, [_vp, _vp, C.c_uint64, _dp, _vp]),
    'psum_vec_scale': (C.c_int, [_vp, C.c_uint64, C.c_double, C.c_double, _vp]),
    'psum_vec_axpy': (C.c_int, [_vp, _vp, C.c_uint64, C.c_double, C.c_double, _vp]),
    'psum_nccl_unique_id': (C.c_int, [_vp]),
    'psum_shard_init': (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    'psum_apply_part': (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _dp, _vp]),
    'psum_shard_terms': (C.c_int, [_vp, C.c_int, C.c_int64, _u64p, _u64p, _dp, _i64p]),
    'psum_apply_sharded': (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, _dp, _vp]),
    'psum_expect_sharded_host': (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, _dp, _vp]),
    'psum_allreduce_sum': (C.c_int, [_vp, _dp, C.c_int]),
    'psum_test_eigh': (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp]),
    'psum_test_schur': (C.c_int, [C.c_int, _dp, _dp, _dp]),
    'psum_test_arnoldi': (C.c_int, [C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, _dp, _dp, _dp,
                                    C.POINTER(C.c_int)]),
}

_lib = None


class PsumError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libpsum error %d: %s' % (code, msg))
        self.code = code


def lib():
    """Loads libpsum.so (built in-tree by qiskit_aqua_b200/build.py or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s not found: run `python -m qiskit_aqua_b200.build` (needs nvcc); '
                              'this package has no CPU fallback' % LIB_PATH)
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != PSUM_OK:
        raise PsumError(rc, lib().psum_last_error().decode('utf-8', 'replace'))
