"""ctypes loader of oracle/psum_oracle.c.  TEST INFRASTRUCTURE ONLY (see pauli_oracle.py header)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libpsum_oracle.so')
_lib = None
_u64p = C.POINTER(C.c_uint64)
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def build(force=False):
    src = os.path.join(HERE, 'psum_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        # -march=native is avoided: the .so travels from the build container to the GPU box
        subprocess.check_call(['gcc', '-O3', '-fopenmp', '-shared', '-fPIC', '-o', LIB, src])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.psum_oracle_threads.restype = C.c_int
        _lib.psum_oracle_set_threads.argtypes = [C.c_int]
        _lib.psum_oracle_state.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _dp]
        _lib.psum_oracle_apply.argtypes = [C.c_int, C.c_int64, _u64p, _u64p, _u8p, _dp, _dp, _dp,
                                           C.c_uint64, C.c_uint64]
        _lib.psum_oracle_apply_rows_counter.argtypes = [C.c_int64, _u64p, _u64p, _u8p, _dp,
                                                        C.c_uint64, C.c_double, _u64p, C.c_int64, _dp]
        _lib.psum_oracle_expect.argtypes = [C.c_int, C.c_int64, _u64p, _u64p, _u8p, _dp, _dp, _dp]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _terms(x, z, ypow, coeff):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    ypow = np.ascontiguousarray(ypow, dtype=np.uint8)
    coeff = np.ascontiguousarray(coeff, dtype=np.complex128)
    return x, z, ypow, coeff


def threads():
    return lib().psum_oracle_threads()


def use_all_cores():
    """OpenMP thread count := cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    lib().psum_oracle_set_threads(n)
    return threads()


def counter_state(seed, first, count):
    out = np.empty(count, dtype=np.complex128)
    lib().psum_oracle_state(seed, first, count, _p(out.view(np.float64), _dp))
    return out


def apply(n, x, z, ypow, coeff, psi, row_begin=0, row_end=None):
    x, z, ypow, coeff = _terms(x, z, ypow, coeff)
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    row_end = (1 << n) if row_end is None else row_end
    y = np.empty(row_end - row_begin, dtype=np.complex128)
    lib().psum_oracle_apply(n, len(x), _p(x, _u64p), _p(z, _u64p), _p(ypow, _u8p),
                            _p(coeff.view(np.float64), _dp), _p(psi.view(np.float64), _dp),
                            _p(y.view(np.float64), _dp), row_begin, row_end)
    return y


def apply_rows_counter(x, z, ypow, coeff, seed, scale, rows):
    x, z, ypow, coeff = _terms(x, z, ypow, coeff)
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    y = np.empty(len(rows), dtype=np.complex128)
    lib().psum_oracle_apply_rows_counter(len(x), _p(x, _u64p), _p(z, _u64p), _p(ypow, _u8p),
                                         _p(coeff.view(np.float64), _dp), seed, scale,
                                         _p(rows, _u64p), len(rows), _p(y.view(np.float64), _dp))
    return y


def expect(n, x, z, ypow, coeff, psi):
    x, z, ypow, coeff = _terms(x, z, ypow, coeff)
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    out = np.zeros(2)
    lib().psum_oracle_expect(n, len(x), _p(x, _u64p), _p(z, _u64p), _p(ypow, _u8p),
                             _p(coeff.view(np.float64), _dp), _p(psi.view(np.float64), _dp),
                             _p(out, _dp))
    return complex(out[0], out[1])
