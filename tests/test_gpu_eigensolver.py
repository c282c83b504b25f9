"""Device Lanczos / diagonal branch against scipy and the reference's golden spectrum (GPU)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co            # noqa: E402
from oracle import pauli_oracle as po        # noqa: E402
from qiskit_aqua_b200 import PauliSumEngine, workloads as W  # noqa: E402

H2 = [(-1.052373245772859, 'II'), (0.39793742484318045, 'ZI'), (-0.39793742484318045, 'IZ'),
      (-0.01128010425623538, 'ZZ'), (0.18093119978423156, 'XX')]


def test_g1_h2_spectrum_lanczos():
    eng = PauliSumEngine.from_labels([l for _, l in H2], [w for w, _ in H2])
    r = eng.lanczos(k=1)
    assert abs(r['evals'][0] - (-1.85727503)) < 1e-7
    r = eng.lanczos(k=4, want_residuals=True)
    np.testing.assert_array_almost_equal(r['evals'], [-1.85727503, -1.24458455, -0.88272215, -0.22491125])
    assert r['resid'].max() < 1e-10


@pytest.mark.parametrize('n,k', [(6, 3), (10, 4), (12, 2)])
def test_lanczos_vs_dense(n, k):
    _, x, z, y, c = W.random_dense_paulis(n, 60, seed=n, complex_coeffs=False)
    eng = PauliSumEngine(n, x, z, y, c)
    N = 1 << n
    H = np.stack([co.apply(n, x, z, y, c, np.eye(N, dtype=complex)[j]) for j in range(N)], axis=1)
    ev = np.linalg.eigvalsh(H)
    r = eng.lanczos(k=k, tol=1e-12, want_residuals=True)
    np.testing.assert_allclose(r['evals'], ev[:k], rtol=0, atol=1e-10 * np.abs(ev).max())
    assert r['resid'].max() < 1e-8 * np.abs(ev).max()
    vecs = r['evecs'].cpu().numpy()
    np.testing.assert_allclose(np.abs(vecs @ vecs.conj().T), np.eye(k), atol=1e-9)


def test_lanczos_three_vector_mode_matches():
    n = 12
    _, x, z, y, c = W.two_local_random(n, 150, seed=8)
    eng = PauliSumEngine(n, x, z, y, c)
    full = eng.lanczos(k=1, tol=1e-12)
    lowmem = eng.lanczos(k=1, tol=1e-12, mem_budget_bytes=5 * 16 * (1 << n), want_residuals=True)
    assert abs(full['evals'][0] - lowmem['evals'][0]) <= 1e-10 * abs(full['evals'][0])
    v = lowmem['evecs'][0]
    hv = eng.apply(v)
    assert torch.linalg.norm(hv - lowmem['evals'][0] * v).item() < 1e-7


def test_lanczos_rejects_non_hermitian():
    eng = PauliSumEngine.from_labels(['XZ', 'ZZ'], [1.0 + 0.5j, 1.0])
    with pytest.raises(Exception) as ei:
        eng.lanczos(k=1)
    assert 'hermitian' in str(ei.value).lower()


def test_config1_maxcut_diagonal_branch():
    # BASELINE config 1: max_cut.get_operator on a seeded 16-node graph -> diagonal branch
    w = po.random_graph(16, weight_range=10, edge_prob=0.5, negative_weight=True, seed=17)
    terms, shift = po.max_cut_operator(w)
    x, z, y, c = po.pack_terms(terms)
    eng = PauliSumEngine(16, x, z, y, c)
    assert eng.info()['is_diagonal'] == 1
    vals, idx = eng.diag_kmin(4)
    diag = po.mf_diag(16, z, c).real
    np.testing.assert_allclose(vals, np.sort(diag)[:4], rtol=0, atol=1e-12)
    for v, i in zip(vals, idx):
        assert diag[int(i)] == pytest.approx(v, abs=1e-12)


def test_config3_molecular_k4():
    # BASELINE config 3 stand-in that fits the oracle: G11 LiH 12-qubit JW operator, k = 4
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, 'golden', 'chemistry_integrals.json')))['G11_lih_npz']
    terms = po.jordan_wigner(po.onee_to_spin(np.array(g['mo_onee'])), po.twoe_to_spin(np.array(g['mo_eri'])))
    x, z, y, c = po.pack_terms(terms)
    eng = PauliSumEngine(12, x, z, y, c)
    r = eng.lanczos(k=4, tol=1e-12, want_residuals=True)
    import scipy.sparse.linalg as spla
    N = 1 << 12
    op = spla.LinearOperator((N, N), dtype=np.complex128,
                             matvec=lambda v: co.apply(12, x, z, y, c, v.astype(complex)))
    ev = np.sort(spla.eigsh(op, k=6, which='SA', return_eigenvectors=False, tol=1e-13))
    # the molecular spectrum has exact degeneracies; compare the distinct lowest values
    assert abs(r['evals'][0] - ev[0]) < 1e-9
    assert abs(round(r['evals'][0] + g['nuclear_repulsion'], 3) - g['expect_total_energy_3dp']) < 1.5e-3
    assert r['resid'].max() < 1e-7


def test_lanczos_finds_degenerate_copies():
    """Exact degeneracies: a single-vector Krylov space sees each eigenvalue once; the probe
    restarts must recover the multiplicities the reference's dense/ARPACK path reports.
    Ferromagnetic Heisenberg ring, 6 sites: spectrum -6 (x7), -4 (x...), ..."""
    n = 6
    labels, coeffs = [], []
    for i in range(n):
        for P in 'XYZ':
            lab = ['I'] * n
            lab[i] = P
            lab[(i + 1) % n] = P
            labels.append(''.join(lab))
            coeffs.append(-1.0)
    eng = PauliSumEngine.from_labels(labels, coeffs)
    terms = list(zip(coeffs, labels))
    ev = np.linalg.eigvalsh(po.ref_to_spmatrix(terms).toarray())
    assert abs(ev[0] - ev[6]) < 1e-10 and ev[7] - ev[6] > 1.0          # 7-fold ground multiplet
    r = eng.lanczos(k=9, tol=1e-12, want_residuals=True)
    np.testing.assert_allclose(r['evals'], ev[:9], rtol=0, atol=1e-9)
    vecs = r['evecs'].cpu().numpy()
    np.testing.assert_allclose(np.abs(vecs @ vecs.conj().T), np.eye(9), atol=1e-8)
    assert r['resid'].max() < 1e-8
    eng.set_option('lanczos_probe', 0)
    r0 = eng.lanczos(k=3, tol=1e-12)
    assert abs(r0['evals'][0] - ev[0]) < 1e-9           # the lowest value never needs the probe


def test_filter_criterion_on_a_non_diagonal_operator_takes_the_dense_branch():
    """filter_criterion makes the reference ask for all 2^n pairs (numpy_eigen_solver.py:249-251 ->
    np.linalg.eig, :171-173).  8 qubits: more pairs than the Lanczos basis holds (used to
    read out of bounds); the dense device branch returns the full sorted spectrum."""
    from qiskit_aqua_b200 import NumPyEigensolver, WeightedPauliOperator, Pauli, workloads as W
    n, x, z, y, c = W.two_local_random(8, 60, seed=8)
    op = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(xx), int(zz), n)] for xx, zz, cc in zip(x, z, c)])
    dense = po.ref_to_spmatrix([(complex(cc), po.masks_to_label(int(xx), int(zz), n)) for xx, zz, cc in zip(x, z, c)]).toarray()
    ref = np.sort(np.linalg.eigvalsh(dense))
    thr = ref[5] - 1e-9
    res = NumPyEigensolver(op, k=3, filter_criterion=lambda v, e, aux: e.real >= thr).run()
    np.testing.assert_allclose(np.real(res.eigenvalues), ref[5:8], atol=1e-9)
    full = NumPyEigensolver(op, k=1 << n).run()
    np.testing.assert_allclose(np.real(full.eigenvalues), ref, atol=1e-9)
    vec = full.eigenstates[0].primitive.cpu().numpy()
    assert np.linalg.norm(dense @ vec - ref[0] * vec) < 1e-8
    # k beyond the Lanczos basis but below 2^n - 1
    some = NumPyEigensolver(op, k=150).run()
    np.testing.assert_allclose(np.real(some.eigenvalues), ref[:150], atol=1e-9)


def test_lanczos_rejects_k_beyond_its_basis():
    from qiskit_aqua_b200 import PauliSumEngine, PsumError, workloads as W
    n, x, z, y, c = W.two_local_random(10, 80, seed=2)
    eng = PauliSumEngine(n, x, z, y, c)
    with pytest.raises(PsumError):
        eng.lanczos(k=200)
    with pytest.raises(PsumError):
        eng.lanczos(k=1 << n)


def test_non_hermitian_sum_small_uses_dense_eig():
    """Complex weights: the reference's eigs(which='SR') is an Arnoldi method and accepts them; tiny sums
    take the dense branch on the device, larger ones the device Arnoldi (psum_arnoldi)."""
    from qiskit_aqua_b200 import NumPyEigensolver, WeightedPauliOperator, Pauli, workloads as W
    n, x, z, y, c = W.random_dense_paulis(5, 40, seed=11, complex_coeffs=True)
    op = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(xx), int(zz), n)] for xx, zz, cc in zip(x, z, c)])
    dense = po.ref_to_spmatrix([(complex(cc), po.masks_to_label(int(xx), int(zz), n)) for xx, zz, cc in zip(x, z, c)]).toarray()
    np.testing.assert_allclose(op.to_opflow().to_matrix(), dense, atol=1e-13)
    ref = np.linalg.eigvals(dense)
    ref = ref[np.argsort(ref)]
    res = NumPyEigensolver(op, k=4).run()
    np.testing.assert_allclose(res.eigenvalues, ref[:4], atol=1e-9)
    # 15 qubits: beyond the dense branch -> Krylov-Schur Arnoldi on the matvec; checked by explicit residuals
    n2, x, z, y, c = W.random_dense_paulis(15, 30, seed=12, complex_coeffs=True)
    big = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(xx), int(zz), n2)] for xx, zz, cc in zip(x, z, c)])
    solver = NumPyEigensolver(big, k=2)
    res = solver.run()
    ev = np.asarray(res.eigenvalues)
    assert ev.dtype == np.complex128 and ev[0].real <= ev[1].real
    eng = big.engine
    scale = np.abs(c).sum()
    for i in range(2):
        v = solver._ret['eigvecs'][i]
        assert abs(torch.linalg.norm(v).item() - 1.0) < 1e-10
        assert torch.linalg.norm(eng.apply(v) - complex(ev[i]) * v).item() < 1e-8 * scale


@pytest.mark.parametrize('n,k,T', [(8, 3, 40), (10, 4, 60), (11, 1, 25)])
def test_arnoldi_vs_dense_eig(n, k, T):
    """psum_arnoldi == the k eigenvalues of smallest real part of the dense matrix (what
    eigs(which='SR') + argsort returns, numpy_eigen_solver.py:175-180), non-Hermitian sums."""
    _, x, z, y, c = W.random_dense_paulis(n, T, seed=100 + n, complex_coeffs=True)
    eng = PauliSumEngine(n, x, z, y, c)
    assert not eng.info()['is_hermitian']
    N = 1 << n
    H = np.stack([co.apply(n, x, z, y, c, np.eye(N, dtype=complex)[j]) for j in range(N)], axis=1)
    ref = np.linalg.eigvals(H)
    ref = ref[np.lexsort((ref.imag, ref.real))]
    r = eng.arnoldi(k=k, tol=1e-12, want_residuals=True)
    scale = np.abs(ref).max()
    np.testing.assert_allclose(r['evals'], ref[:k], rtol=0, atol=1e-9 * scale)
    assert r['resid'].max() < 1e-8 * scale
    vecs = r['evecs'].cpu().numpy()
    for i in range(k):
        assert np.linalg.norm(H @ vecs[i] - r['evals'][i] * vecs[i]) < 1e-8 * scale


def test_arnoldi_equals_lanczos_on_a_hermitian_sum():
    n = 12
    _, x, z, y, c = W.two_local_random(n, 150, seed=8)
    eng = PauliSumEngine(n, x, z, y, c)
    a = eng.arnoldi(k=3, tol=1e-12, want_residuals=True)
    l = eng.lanczos(k=3, tol=1e-12)
    np.testing.assert_allclose(a['evals'].real, l['evals'], rtol=0, atol=1e-9 * np.abs(l['evals']).max())
    assert np.abs(a['evals'].imag).max() < 1e-9
    assert a['resid'].max() < 1e-7


def test_diagonal_operator_with_filter_uses_one_device_sort():
    """MaxCut-style diagonal operator with a filter: the reference sorts the diagonal once (:163-169);
    here one device sort replaces k = 2^n min-sweeps."""
    from qiskit_aqua_b200 import NumPyEigensolver, WeightedPauliOperator, Pauli
    rng = np.random.default_rng(3)
    n = 12
    paulis = []
    for _ in range(40):
        i, j = sorted(rng.choice(n, 2, replace=False).tolist())
        paulis.append([float(rng.integers(1, 5)), Pauli.from_masks(0, (1 << i) | (1 << j), n)])
    op = WeightedPauliOperator(paulis=paulis)
    diag = np.zeros(1 << n)
    idx = np.arange(1 << n)
    for w, p in op.paulis:
        diag += np.real(w) * (1.0 - 2.0 * (np.bitwise_count((idx & p.z_mask).astype(np.uint64)) & 1).astype(np.float64))
    ref = np.sort(diag)
    res = NumPyEigensolver(op, k=5, filter_criterion=lambda v, e, aux: e.real > ref[99]).run()
    want = ref[ref > ref[99]][:5]
    np.testing.assert_allclose(np.real(res.eigenvalues), want, atol=1e-12)
    # complex weights on a diagonal operator: np.sort orders by (real, imag)
    opc = WeightedPauliOperator(paulis=[[w * (1 + 0.5j), p] for w, p in paulis])
    resc = NumPyEigensolver(opc, k=3).run()
    refc = np.sort(diag * (1 + 0.5j))[:3]
    np.testing.assert_allclose(resc.eigenvalues, refc, atol=1e-12)


def test_config4_lanczos_ground_state_tol_1e12():
    """BASELINE config 4: 30-qubit TFIM + Heisenberg + 3-local Hamiltonian, 10 000 terms,
    on-device Lanczos ground state at tol = 1e-12.  The explicit residual |H v - E v| and the Rayleigh
    quotient of the returned vector are recomputed with independent kernel launches.  (About 200
    matvecs of ~0.35 s plus the second pass for the vector: ~2.5 minutes on a B200; smaller devices
    run the same operator family at fewer qubits.)"""
    from qiskit_aqua_b200.engine import vec_dot
    free_b, _ = torch.cuda.mem_get_info()
    n = 30 if free_b > 120 * 2 ** 30 else 26
    pairs = n * (n - 1) // 2
    _, x, z, y, c = W.ising_heisenberg(n, n_triples=max(0, 10000 - 3 * pairs - n), seed=30)
    eng = PauliSumEngine(n, x, z, y, c)
    r = eng.lanczos(k=1, tol=1e-12, max_iter=600, seed=4, want_vectors=True)
    v = r['evecs'][0]
    e0 = float(r['evals'][0])
    hv = eng.apply(v)
    rayleigh = vec_dot(v, hv).real
    hv.add_(v, alpha=-e0)
    resid = float(torch.linalg.vector_norm(hv).item())
    assert abs(float(torch.linalg.vector_norm(v).item()) - 1.0) < 1e-12
    assert resid <= 1e-8, resid
    assert abs(rayleigh - e0) <= 1e-10 * abs(e0), (rayleigh, e0)
    if n == 30:
        assert abs(e0 - (-35.90543862024841)) < 1e-8       # profiles/config4_lanczos_r2_q30_tol1e-12.json
