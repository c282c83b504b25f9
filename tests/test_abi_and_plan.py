"""CPU-only: the C-ABI library loads and exports every symbol include/psum.h declares, the
host-side packing / planning logic, and the host eigen-solvers.  No compute calls (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pauli_oracle as po
from qiskit_aqua_b200 import PauliSumEngine, PsumError, _lib, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, 'include', 'psum.h')).read()
    names = set(re.findall(r'\b(psum_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 24
    lib = C.CDLL(_lib.LIB_PATH)
    for nm in sorted(names):
        assert hasattr(lib, nm), 'libpsum.so lacks %s' % nm
    assert set(_lib.SIGNATURES) == names          # the ctypes table binds exactly the header


def test_no_cpu_fallback_without_device():
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    eng = PauliSumEngine.from_labels(['XX', 'ZI'], [1.0, 0.5])
    with pytest.raises(PsumError) as ei:
        eng.expect(np.ones(4, dtype=complex))
    assert ei.value.code == _lib.ERR_CUDA


def test_merge_and_flags():
    eng = PauliSumEngine.from_labels(['XXI', 'XXI', 'ZZI', 'IYY', 'ZZI'], [0.5, 0.25, 1.0, 1j, -1.0])
    info = eng.info()
    assert info['n_terms'] == 2 and info['n_xgroups'] == 2
    assert info['is_hermitian'] == 0 and info['is_diagonal'] == 0
    eng = PauliSumEngine.from_labels(['ZZI', 'IIZ'], [1.0, 2.0])
    assert eng.info()['is_diagonal'] == 1 and eng.info()['is_hermitian'] == 1
    eng = PauliSumEngine.from_labels(['YI', 'XY'], [1.0, -2.0])      # real coeffs on label Paulis
    assert eng.info()['is_hermitian'] == 1


def test_invalid_arguments():
    x = np.array([1 << 5], dtype=np.uint64)
    z = np.zeros(1, dtype=np.uint64)
    with pytest.raises(PsumError):
        PauliSumEngine(3, x, z, np.zeros(1, np.uint8), np.ones(1, complex))      # mask above n
    with pytest.raises(PsumError):
        PauliSumEngine(0, x[:0], z[:0], np.zeros(0, np.uint8), np.ones(0, complex))
    eng = PauliSumEngine.from_labels(['XX'], [1.0])
    with pytest.raises(PsumError):
        eng.set_option('tile_bits', 99)
    with pytest.raises(PsumError):
        eng.set_option('no_such_option', 1)
    with pytest.raises(PsumError):
        eng.shard(0, 3)                                                         # not a power of 2
    with pytest.raises(PsumError):
        eng.set_option('plan_tries', 0)
    with pytest.raises(PsumError):
        eng.set_option('reg_bits', 2)
    eng.set_option('term_cache', 0)
    eng.set_option('plan_tries', 8)


def test_planner_restarts_only_ever_lower_the_cost():
    """More randomised restarts never give a plan with more HBM sweeps (the cheapest plan wins), and the
    kernel-side options do not change the plan."""
    n, x, z, y, c = W.ising_heisenberg(24, n_triples=2000, seed=7)
    counts = []
    for tries in (1, 8, 64):
        eng = PauliSumEngine(n, x, z, y, c, {'plan_tries': tries})
        info = eng.info()
        counts.append(3 * info['n_passes'] + info['n_remote_groups'])
    assert counts[0] >= counts[1] >= counts[2]
    a = PauliSumEngine(n, x, z, y, c, {'term_cache': 0}).info()
    b = PauliSumEngine(n, x, z, y, c, {'term_cache': 1}).info()
    assert a == b


@pytest.mark.parametrize('tile_bits', [11, 12, 13])
def test_plan_covers_every_group_once(tile_bits):
    n, x, z, y, c = W.ising_heisenberg(20, n_triples=600, seed=5)
    eng = PauliSumEngine(n, x, z, y, c, {'tile_bits': tile_bits})
    info = eng.info()
    assert info['tile_bits'] == tile_bits
    tot = rem = 0
    for p in range(info['n_passes']):
        pi = eng.pass_info(p)
        assert bin(pi['free_mask']).count('1') == tile_bits
        assert pi['free_mask'] & 0x7 == 0x7                       # the 3 low qubits are always free (128-byte runs)
        tot += pi['n_groups']
        rem += pi['n_remote_groups']
    assert tot == info['n_xgroups'] == len(set(x.tolist()))
    assert rem == info['n_remote_groups']
    assert info['n_patterns'] <= info['n_component_terms']
    assert info['n_passes'] < info['n_xgroups'] / 4               # tiling pays: few HBM sweeps


def test_headline_plan_reaches_the_pair_covering_bound():
    """The 30-qubit headline operator (bench.py, config C4) holds every XX/YY pair mask: all pairs of the 27
    high qubits must share a free set of 9 high qubits, which needs >= ceil(27/9 * ceil(26/8)) = 12 passes
    (Schoenheim bound).  The planner's randomised restarts reach it; on 8 ranks (33 qubits) the local part
    plus the six exchanged parts take 24."""
    import bench
    n, x, z, y, c = bench.workload(30, 'c4')
    info = PauliSumEngine(n, x, z, y, c).info()
    assert info['n_xgroups'] == 466 and info['n_remote_groups'] == 0
    assert info['n_passes'] == 12
    n, x, z, y, c = bench.workload(33, 'c4')
    eng = PauliSumEngine(n, x, z, y, c)
    eng.shard(0, 8)
    assert eng.info()['n_passes'] <= 24


def test_small_operator_uses_streaming_plan():
    n, x, z, y, c = W.two_local_random(8, 40, seed=1)
    assert PauliSumEngine(n, x, z, y, c).info()['n_passes'] == 0


def test_shard_terms_algebra_world4():
    """Sum over ranks/parts of the sharded terms reproduces H.psi (pure host logic)."""
    n, world = 9, 4
    _, x, z, y, c = W.random_dense_paulis(n, 120, seed=4, complex_coeffs=True)
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    ref = po.mf_apply(n, x, z, y, c, psi)
    n_loc = n - 2
    N_loc = 1 << n_loc
    out = np.zeros_like(ref)
    for rank in range(world):
        eng = PauliSumEngine(n, x, z, y, c)
        eng.shard(rank, world)
        assert eng.info()['n_local_qubits'] == n_loc
        for g in range(world):
            xl, zl, cl = eng.shard_terms(g)
            if len(xl) == 0:
                continue
            src = psi[((rank ^ g) * N_loc):((rank ^ g) * N_loc + N_loc)]
            out[rank * N_loc:(rank + 1) * N_loc] += po.mf_apply(
                n_loc, xl, zl, np.zeros(len(xl), np.uint8), cl, src)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-11 * np.abs(ref).max())


def test_host_eigensolvers():
    lib = _lib.lib()
    rng = np.random.default_rng(3)
    dp = C.POINTER(C.c_double)
    for n in (1, 2, 7, 40):
        A = rng.standard_normal((n, n))
        A = A + A.T
        w = np.zeros(n)
        V = np.zeros((n, n))
        assert lib.psum_test_eigh(0, n, A.ctypes.data_as(dp), w.ctypes.data_as(dp), V.ctypes.data_as(dp)) == 0
        np.testing.assert_allclose(w, np.linalg.eigvalsh(A), atol=1e-12 * max(1, np.abs(A).max() * n))
        np.testing.assert_allclose(A @ V, V * w, atol=1e-11 * n)
        d, e = rng.standard_normal(n), rng.standard_normal(n)
        T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
        de = np.concatenate([d, e])
        assert lib.psum_test_eigh(1, n, de.ctypes.data_as(dp), w.ctypes.data_as(dp), V.ctypes.data_as(dp)) == 0
        np.testing.assert_allclose(w, np.linalg.eigvalsh(T), atol=1e-12 * n)
        np.testing.assert_allclose(T @ V, V * w, atol=1e-11 * n)


def test_host_schur_and_arnoldi_driver():
    """The host half of psum_arnoldi: complex Schur form (sorted by real part) and the Krylov-Schur
    restart logic, run on dense matrices with host vectors (psum_test_schur / psum_test_arnoldi)."""
    lib = _lib.lib()
    rng = np.random.default_rng(5)
    dp = C.POINTER(C.c_double)
    for n in (1, 2, 5, 33, 128):
        for kind in range(3):
            A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            if kind == 1:
                A = A + A.conj().T
            if kind == 2:      # triangular with repeated integer eigenvalues
                A = np.triu(A)
                A[np.diag_indices(n)] = np.round(A.diagonal().real)
            A = np.ascontiguousarray(A)
            T, Z = np.zeros_like(A), np.zeros_like(A)
            assert lib.psum_test_schur(n, A.ctypes.data_as(dp), T.ctypes.data_as(dp), Z.ctypes.data_as(dp)) == 0
            tol = 1e-12 * max(1.0, np.abs(A).max() * n)
            np.testing.assert_allclose(Z @ T @ Z.conj().T, A, atol=tol)
            np.testing.assert_allclose(Z.conj().T @ Z, np.eye(n), atol=1e-12)
            assert np.abs(np.tril(T, -1)).max() == 0.0
            d = T.diagonal()
            assert all((d[i].real, d[i].imag) <= (d[i + 1].real, d[i + 1].imag) for i in range(n - 1))
            np.testing.assert_allclose(np.sort_complex(d), np.sort_complex(np.linalg.eigvals(A)), atol=100 * tol)
    for n, k, m in ((6, 2, 0), (30, 3, 0), (200, 4, 0), (300, 6, 20), (300, 1, 12)):
        for kind in range(3):
            A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            if kind == 1:
                A = (A + A.conj().T) / 2
            if kind == 2:      # Hermitian plus a small anti-Hermitian part
                A = (A + A.conj().T) / 2 + 0.05 * (A - A.conj().T)
            A = np.ascontiguousarray(A)
            ev, vec, rs, it = np.zeros(2 * k), np.zeros((k, n), dtype=np.complex128), np.zeros(k), C.c_int(0)
            rc = lib.psum_test_arnoldi(n, A.ctypes.data_as(dp), k, m, 1e-12, 5000, 1, ev.ctypes.data_as(dp),
                                       vec.ctypes.data_as(dp), rs.ctypes.data_as(dp), C.byref(it))
            assert rc == 0, _lib.last_error() if hasattr(_lib, 'last_error') else rc
            ev = ev[0::2] + 1j * ev[1::2]
            ref = np.linalg.eigvals(A)
            ref = ref[np.lexsort((ref.imag, ref.real))][:k]
            np.testing.assert_allclose(ev, ref, atol=1e-8)
            for i in range(k):
                assert np.linalg.norm(A @ vec[i] - ev[i] * vec[i]) < 1e-8
                assert abs(np.linalg.norm(vec[i]) - 1.0) < 1e-10


def test_c_oracle_matches_numpy_oracle():
    from oracle import c_oracle as co
    n, x, z, y, c = W.random_dense_paulis(13, 80, seed=11, complex_coeffs=True)
    rng = np.random.default_rng(2)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    ref = po.mf_apply(n, x, z, y, c, psi)
    np.testing.assert_allclose(co.apply(n, x, z, y, c, psi), ref, atol=1e-12 * np.abs(ref).max())
    np.testing.assert_allclose(co.apply(n, x, z, y, c, psi, 3, 4099), ref[3:4099], atol=1e-12 * np.abs(ref).max())
    assert abs(co.expect(n, x, z, y, c, psi) - np.vdot(psi, ref)) < 1e-9
    s = co.counter_state(5, 0, 1 << n)
    rows = np.array([0, 17, (1 << n) - 1], dtype=np.uint64)
    np.testing.assert_allclose(co.apply_rows_counter(x, z, y, c, 5, 1.0, rows),
                               po.mf_apply(n, x, z, y, c, s, rows), atol=1e-12 * np.abs(ref).max())


def test_plan_up_to_64_qubits_and_sharded():
    rng = np.random.default_rng(64)
    for n in (40, 64):
        T = 300
        x = np.zeros(T, dtype=np.uint64)
        z = np.zeros(T, dtype=np.uint64)
        for k in range(T):
            for q in rng.choice(n, size=3, replace=False):
                p = rng.integers(0, 3)
                if p < 2:
                    x[k] |= np.uint64(1) << np.uint64(int(q))
                if p > 0:
                    z[k] |= np.uint64(1) << np.uint64(int(q))
        y = np.array([bin(int(a) & int(b)).count('1') & 3 for a, b in zip(x, z)], dtype=np.uint8)
        eng = PauliSumEngine(n, x, z, y, rng.standard_normal(T).astype(complex))
        info = eng.info()
        assert info['n_qubits'] == n and info['n_passes'] >= 1
        assert sum(eng.pass_info(p)['n_groups'] for p in range(info['n_passes'])) == info['n_xgroups']
        eng.shard(5, 8)
        info = eng.info()
        assert info['n_local_qubits'] == n - 3 and 1 <= info['n_global_parts'] <= 7


def test_late_bits_keeps_early_passes_chunk_local():
    n, x, z, y, c = W.ising_heisenberg(24, n_triples=2000, seed=3)
    eng = PauliSumEngine(n, x, z, y, c, {'late_bits': 4})
    info = eng.info()
    masks = [eng.pass_info(p)['free_mask'] for p in range(info['n_passes'])]
    early = [m for m in masks if m < (1 << 20)]
    assert len(early) >= info['n_passes'] // 2            # most passes never touch the top 4 qubits
    k = len(early)
    assert all(m < (1 << 20) for m in masks[:k])           # and they come first in the plan
    assert sum(eng.pass_info(p)['n_groups'] for p in range(info['n_passes'])) == info['n_xgroups']
