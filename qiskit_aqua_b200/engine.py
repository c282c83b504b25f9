"""PauliSumEngine: the packed Pauli sum living behind libpsum.so.

Host side of the drop-in boundary.  A term is packed ONCE into (x_mask, z_mask, ypow, coeff);
state vectors are torch.complex128 CUDA tensors used purely as device-buffer carriers.
Replaces the CSR assembly + spmv of the reference:
  legacy/op_converter.py:103-124, legacy/weighted_pauli_operator.py:605-623,
  primitive_ops/pauli_op.py:151-164, list_ops/list_op.py:320-329.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import PsumError, check, lib

INFO_KEYS = ('n_qubits', 'n_terms', 'n_xgroups', 'n_passes', 'n_remote_groups', 'is_hermitian',
             'is_diagonal', 'n_local_qubits', 'n_patterns', 'n_component_terms', 'tile_bits',
             'n_global_parts', 'n_flat_groups', 'n_group_records', 'n_fast_groups', 'n_bundles', 'n_quad_groups')


def pack_masks(labels_or_masks):
    """list of big-endian labels -> (x_mask u64[T], z_mask u64[T], ypow u8[T]).
    Label convention: leftmost char = qubit n-1 (terra Pauli labels)."""
    xs, zs, ys = [], [], []
    for lab in labels_or_masks:
        x = z = 0
        n = len(lab)
        for pos, ch in enumerate(lab):
            bit = 1 << (n - 1 - pos)
            if ch == 'X':
                x |= bit
            elif ch == 'Z':
                z |= bit
            elif ch == 'Y':
                x |= bit
                z |= bit
            elif ch != 'I':
                raise ValueError('invalid Pauli label %r' % lab)
        xs.append(x)
        zs.append(z)
        ys.append(bin(x & z).count('1') & 3)
    return (np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64),
            np.array(ys, dtype=np.uint8))


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_device_state(state, n_amps=None, device=None):
    """numpy / torch (host or device) -> contiguous complex128 CUDA tensor (H2D copy if needed)."""
    if not torch.cuda.is_available():
        raise PsumError(_lib.ERR_CUDA, 'no CUDA device available; qiskit_aqua_b200 has no CPU fallback')
    if isinstance(state, torch.Tensor):
        t = state
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(state, dtype=np.complex128)))
    if t.dtype != torch.complex128:
        t = t.to(torch.complex128)
    if not t.is_cuda:
        t = t.to(device or 'cuda', non_blocking=True)
    t = t.contiguous().reshape(-1)
    if n_amps is not None and t.numel() != n_amps:
        raise ValueError('state vector has %d amplitudes, expected %d' % (t.numel(), n_amps))
    return t


class PauliSumEngine:
    """Owns a psum_handle_t."""

    def __init__(self, num_qubits, x_mask, z_mask, ypow, coeff, options=None):
        self._h = C.c_void_p()
        self.num_qubits = int(num_qubits)
        x_mask = np.ascontiguousarray(x_mask, dtype=np.uint64)
        z_mask = np.ascontiguousarray(z_mask, dtype=np.uint64)
        ypow = np.ascontiguousarray(ypow, dtype=np.uint8)
        coeff = np.ascontiguousarray(coeff, dtype=np.complex128)
        T = x_mask.shape[0]
        if not (z_mask.shape[0] == ypow.shape[0] == coeff.shape[0] == T):
            raise ValueError('term arrays must have equal length')
        cview = coeff.view(np.float64)
        check(lib().psum_create(self.num_qubits, T,
                                x_mask.ctypes.data_as(C.POINTER(C.c_uint64)),
                                z_mask.ctypes.data_as(C.POINTER(C.c_uint64)),
                                ypow.ctypes.data_as(C.POINTER(C.c_uint8)),
                                cview.ctypes.data_as(C.POINTER(C.c_double)),
                                C.byref(self._h)))
        self.rank, self.world = 0, 1
        for k, v in (options or {}).items():
            self.set_option(k, v)

    @classmethod
    def from_labels(cls, labels, coeffs, options=None):
        labels = list(labels)
        if not labels:
            raise ValueError('need at least one label (use num_qubits + empty arrays otherwise)')
        x, z, y = pack_masks(labels)
        return cls(len(labels[0]), x, z, y, np.asarray(coeffs, dtype=np.complex128), options)

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            lib().psum_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    # a packed operator is immutable: copies of its owner share it (one handle, freed once)
    def __copy__(self):
        return self

    def __deepcopy__(self, memo):
        return self

    # -- introspection ------------------------------------------------------------------------
    def set_option(self, key, value):
        check(lib().psum_set_option(self._h, key.encode(), int(value)))

    def info(self):
        buf = (C.c_int64 * 17)()
        check(lib().psum_info(self._h, buf, 17))
        return dict(zip(INFO_KEYS, [int(v) for v in buf]))

    def pass_info(self, p, g=0):
        fm, ng, nr, npat = C.c_uint64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().psum_pass_info(self._h, g, p, C.byref(fm), C.byref(ng), C.byref(nr), C.byref(npat)))
        return {'free_mask': fm.value, 'n_groups': ng.value, 'n_remote_groups': nr.value,
                'n_patterns': npat.value}

    def plan_floors(self, store=True):
        """Lower bounds of one H.psi on this rank from the plan: dict(dram_bytes, smem_wavefronts,
        group_records, partner_load_sets) (psum_plan_floors)."""
        out = (C.c_double * 4)()
        check(lib().psum_plan_floors(self._h, 1 if store else 0, out))
        return {'dram_bytes': out[0], 'smem_wavefronts': out[1], 'group_records': int(out[2]),
                'partner_load_sets': int(out[3])}

    def exchange_stats(self):
        """(exchanges, bytes per exchange, ms on the communication stream, span ms) of the last sharded
        call made with option profile_exchange = 1."""
        out = (C.c_double * 4)()
        check(lib().psum_exchange_stats(self._h, out))
        return [float(v) for v in out]

    @property
    def n_local_amps(self):
        return 1 << (self.num_qubits - (self.world.bit_length() - 1))

    # -- the hot path ---------------------------------------------------------------------------
    def apply(self, psi, out=None, want_expect=False):
        """y = H psi (device tensor).  Returns y or (y, <psi|H|psi>)."""
        psi = as_device_state(psi, self.n_local_amps)
        if out is None:
            out = torch.empty_like(psi)
        ex = (C.c_double * 2)()
        check(lib().psum_apply(self._h, C.c_void_p(psi.data_ptr()), C.c_void_p(out.data_ptr()),
                               ex if want_expect else None, _stream_ptr()))
        return (out, complex(ex[0], ex[1])) if want_expect else out

    def expect(self, psi):
        """<psi|H|psi> as a python complex (legacy/weighted_pauli_operator.py:621-622)."""
        psi = as_device_state(psi, self.n_local_amps)
        ex = (C.c_double * 2)()
        check(lib().psum_expect(self._h, C.c_void_p(psi.data_ptr()), ex, _stream_ptr()))
        return complex(ex[0], ex[1])

    def expect_host(self, psi_host, scratch=None):
        """<psi|H|psi> for a HOST state (numpy array or CPU torch tensor, ideally pinned): the H2D
        copy is pipelined with the chunk-local passes (psum_expect_host).  `scratch` is an optional
        complex128 CUDA tensor of 2^n amplitudes that receives the copy."""
        if not torch.cuda.is_available():
            raise PsumError(_lib.ERR_CUDA, 'no CUDA device available; qiskit_aqua_b200 has no CPU fallback')
        if isinstance(psi_host, torch.Tensor):
            t = psi_host
            if t.is_cuda:
                return self.expect(t)
            if t.dtype != torch.complex128 or not t.is_contiguous():
                t = t.to(torch.complex128).contiguous()
            host_ptr, keep = t.data_ptr(), t
            n = t.numel()
        else:
            a = np.ascontiguousarray(np.asarray(psi_host, dtype=np.complex128)).reshape(-1)
            host_ptr, keep = a.ctypes.data, a
            n = a.size
        if n != self.n_local_amps:
            raise ValueError('state vector has %d amplitudes, expected %d' % (n, self.n_local_amps))
        if scratch is None:
            scratch = torch.empty(n, dtype=torch.complex128, device='cuda')
        ex = (C.c_double * 2)()
        check(lib().psum_expect_host(self._h, C.c_void_p(host_ptr), C.c_void_p(scratch.data_ptr()), ex,
                                     _stream_ptr()))
        del keep
        return complex(ex[0], ex[1])

    def bra_apply(self, phi, psi):
        phi = as_device_state(phi, self.n_local_amps)
        psi = as_device_state(psi, self.n_local_amps)
        ex = (C.c_double * 2)()
        check(lib().psum_bra_apply(self._h, C.c_void_p(phi.data_ptr()), C.c_void_p(psi.data_ptr()),
                                   ex, _stream_ptr()))
        return complex(ex[0], ex[1])

    def expect_batch(self, psis):
        if isinstance(psis, (list, tuple)):
            psis = torch.stack([as_device_state(p, self.n_local_amps) for p in psis])
        t = as_device_state(psis)
        batch = t.numel() // self.n_local_amps
        if batch * self.n_local_amps != t.numel():
            raise ValueError('batch of state vectors has the wrong size')
        ex = (C.c_double * (2 * max(batch, 1)))()
        check(lib().psum_expect_batch(self._h, batch, C.c_void_p(t.data_ptr()), ex, _stream_ptr()))
        return np.array([complex(ex[2 * i], ex[2 * i + 1]) for i in range(batch)])

    def to_dense(self):
        """Dense matrix of the operator as a [2^n, 2^n] complex128 CUDA tensor (small n only)."""
        if not torch.cuda.is_available():
            raise PsumError(_lib.ERR_CUDA, 'no CUDA device available; qiskit_aqua_b200 has no CPU fallback')
        N = self.n_local_amps
        out = torch.empty((N, N), dtype=torch.complex128, device='cuda')
        check(lib().psum_to_dense(self._h, C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out

    # -- diagonal operators -----------------------------------------------------------------------
    def diag(self):
        d = torch.empty(self.n_local_amps, dtype=torch.float64, device='cuda')
        check(lib().psum_diag(self._h, C.c_void_p(d.data_ptr()), _stream_ptr()))
        return d

    def diag_kmin(self, k):
        vals = (C.c_double * k)()
        idx = (C.c_uint64 * k)()
        check(lib().psum_diag_kmin(self._h, k, vals, idx, _stream_ptr()))
        return np.array(vals[:], dtype=np.float64), np.array(idx[:], dtype=np.uint64)

    # -- eigensolver --------------------------------------------------------------------------------
    def lanczos(self, k=1, tol=1e-12, max_iter=2000, seed=1, mem_budget_bytes=0,
                want_vectors=True, want_residuals=False):
        """k lowest eigenpairs (numpy_eigen_solver.py:175-176).  Returns dict(evals, evecs, resid,
        iters); evecs is a [k, N_loc] complex128 CUDA tensor."""
        if not torch.cuda.is_available():
            raise PsumError(_lib.ERR_CUDA, 'no CUDA device available')
        ev = (C.c_double * k)()
        rs = (C.c_double * k)()
        it = C.c_int(0)
        vecs = torch.empty((k, self.n_local_amps), dtype=torch.complex128, device='cuda') \
            if want_vectors else None
        check(lib().psum_lanczos(self._h, k, float(tol), int(max_iter), int(seed),
                                 int(mem_budget_bytes), ev,
                                 C.c_void_p(vecs.data_ptr()) if want_vectors else None,
                                 rs if want_residuals else None, C.byref(it), _stream_ptr()))
        return {'evals': np.array(ev[:]), 'evecs': vecs,
                'resid': np.array(rs[:]) if want_residuals else None, 'iters': it.value}

    def arnoldi(self, k=1, tol=1e-12, max_iter=3000, seed=1, mem_budget_bytes=0,
                want_vectors=True, want_residuals=False):
        """k eigenpairs of smallest real part of a general (non-Hermitian) Pauli sum -- eigs(which='SR') at
        numpy_eigen_solver.py:175-176 for complex label coefficients.  Returns dict(evals complex128[k]
        ascending by (re, im), evecs [k, N_loc] CUDA tensor, resid, iters)."""
        if not torch.cuda.is_available():
            raise PsumError(_lib.ERR_CUDA, 'no CUDA device available')
        ev = (C.c_double * (2 * k))()
        rs = (C.c_double * k)()
        it = C.c_int(0)
        vecs = torch.empty((k, self.n_local_amps), dtype=torch.complex128, device='cuda') \
            if want_vectors else None
        check(lib().psum_arnoldi(self._h, k, float(tol), int(max_iter), int(seed),
                                 int(mem_budget_bytes), ev,
                                 C.c_void_p(vecs.data_ptr()) if want_vectors else None,
                                 rs if want_residuals else None, C.byref(it), _stream_ptr()))
        e = np.array(ev[:])
        return {'evals': e[0::2] + 1j * e[1::2], 'evecs': vecs,
                'resid': np.array(rs[:]) if want_residuals else None, 'iters': it.value}

    # -- time evolution ------------------------------------------------------------------------------
    def evolve(self, psi, t, tol=1e-10, krylov_dim=0, out=None):
        """exp(-i t H) psi on the device (legacy/matrix_operator.py:354-355 without the matrix).
        Returns (state tensor, info dict)."""
        psi = as_device_state(psi, self.n_local_amps)
        if out is None:
            out = torch.empty_like(psi)
        steps, mv = C.c_int(0), C.c_int(0)
        check(lib().psum_evolve(self._h, C.c_void_p(psi.data_ptr()), C.c_void_p(out.data_ptr()), float(t),
                                float(tol), int(krylov_dim), C.byref(steps), C.byref(mv), _stream_ptr()))
        return out, {'steps': steps.value, 'matvecs': mv.value}

    # -- sharding -------------------------------------------------------------------------------------
    def shard(self, rank, world, unique_id=None):
        buf = None
        if unique_id is not None:
            buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(lib().psum_shard_init(self._h, rank, world, buf))
        self.rank, self.world = rank, world

    def shard_terms(self, g):
        n = C.c_int64()
        check(lib().psum_shard_terms(self._h, g, 0, None, None, None, C.byref(n)))
        x = np.zeros(n.value, dtype=np.uint64)
        z = np.zeros(n.value, dtype=np.uint64)
        c = np.zeros(n.value, dtype=np.complex128)
        if n.value:
            check(lib().psum_shard_terms(self._h, g, n.value, x.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         z.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         c.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                         C.byref(n)))
        return x, z, c

    def apply_part(self, g, src, bra, y, accumulate, want_expect=False):
        ex = (C.c_double * 2)(0.0, 0.0)
        check(lib().psum_apply_part(self._h, g, C.c_void_p(src.data_ptr()),
                                    C.c_void_p(bra.data_ptr()) if bra is not None else None,
                                    C.c_void_p(y.data_ptr()) if y is not None else None,
                                    1 if accumulate else 0, ex if want_expect else None, _stream_ptr()))
        return complex(ex[0], ex[1])

    def expect_sharded_host(self, psi_host, scratch, recv=None):
        """All-reduced <psi|H|psi> for this rank's shard in HOST memory (psum_expect_sharded_host);
        `scratch` is the device buffer that receives the shard."""
        ptr, keep, n = _host_ptr(psi_host)
        if n != self.n_local_amps or scratch.numel() < n:
            raise ValueError('shard has %d amplitudes, expected %d' % (n, self.n_local_amps))
        ex = (C.c_double * 2)()
        check(lib().psum_expect_sharded_host(self._h, C.c_void_p(ptr), C.c_void_p(scratch.data_ptr()),
                                             C.c_void_p(recv.data_ptr()) if recv is not None else None,
                                             recv.numel() if recv is not None else 0, ex, _stream_ptr()))
        del keep
        return complex(ex[0], ex[1])

    def apply_sharded(self, psi, out=None, recv=None, want_expect=False, store=True):
        psi = as_device_state(psi, self.n_local_amps)
        if out is None and store:
            out = torch.empty_like(psi)
        ex = (C.c_double * 2)()
        check(lib().psum_apply_sharded(self._h, C.c_void_p(psi.data_ptr()),
                                       C.c_void_p(out.data_ptr()) if store else None,
                                       C.c_void_p(recv.data_ptr()) if recv is not None else None,
                                       recv.numel() if recv is not None else 0,
                                       ex if want_expect else None, _stream_ptr()))
        return (out, complex(ex[0], ex[1])) if want_expect else out


def _host_ptr(psi_host):
    if isinstance(psi_host, torch.Tensor):
        t = psi_host
        if t.dtype != torch.complex128 or not t.is_contiguous():
            t = t.to(torch.complex128).contiguous()
        return t.data_ptr(), t, t.numel()
    a = np.ascontiguousarray(np.asarray(psi_host, dtype=np.complex128)).reshape(-1)
    return a.ctypes.data, a, a.size


def expect_multi(engines, psi):
    """[<psi|H_j|psi>] for several packed operators against one device state with a single host
    synchronisation (psum_expect_multi)."""
    engines = list(engines)
    if not engines:
        return np.zeros(0, dtype=np.complex128)
    psi = as_device_state(psi, engines[0].n_local_amps)
    hs = (C.c_void_p * len(engines))(*[e._h for e in engines])
    out = (C.c_double * (2 * len(engines)))()
    check(lib().psum_expect_multi(hs, len(engines), C.c_void_p(psi.data_ptr()), out, _stream_ptr()))
    return np.array([complex(out[2 * j], out[2 * j + 1]) for j in range(len(engines))])


def pauli_decompose_device(matrix):
    """W[x, z] = weight(x, z) * (-i)^{popc(x & z)} of a dense 2^n x 2^n matrix (numpy or CUDA tensor),
    computed on the device (psum_pauli_decompose); returns a [2^n, 2^n] complex128 CUDA tensor."""
    if not torch.cuda.is_available():
        raise PsumError(_lib.ERR_CUDA, 'no CUDA device available; qiskit_aqua_b200 has no CPU fallback')
    m = matrix if isinstance(matrix, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(matrix, dtype=np.complex128))
    m = m.to(device='cuda', dtype=torch.complex128).contiguous()
    dim = m.shape[0]
    n = dim.bit_length() - 1
    out = torch.empty_like(m)
    check(lib().psum_pauli_decompose(C.c_void_p(m.data_ptr()), n, C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def nccl_unique_id():
    buf = (C.c_char * 128)()
    check(lib().psum_nccl_unique_id(buf))
    return bytes(buf.raw)


def fill_state(n_amps, seed, first_index=0, out=None):
    """Counter-based synthetic state on the device (twin: oracle.psum_oracle.counter_state)."""
    if out is None:
        out = torch.empty(n_amps, dtype=torch.complex128, device='cuda')
    check(lib().psum_fill_state(C.c_void_p(out.data_ptr()), n_amps, first_index, seed, _stream_ptr()))
    return out


def vec_dot(a, b):
    ex = (C.c_double * 2)()
    check(lib().psum_vec_dot(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), a.numel(), ex,
                             _stream_ptr()))
    return complex(ex[0], ex[1])


def vec_scale(a, s):
    s = complex(s)
    check(lib().psum_vec_scale(C.c_void_p(a.data_ptr()), a.numel(), s.real, s.imag, _stream_ptr()))
    return a
