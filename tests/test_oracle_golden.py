"""Pins the CPU oracle (oracle/pauli_oracle.py) against the reference's own golden vectors
(SURVEY.md section 8c G1..G13).  CPU only."""
import itertools
import json
import os

import numpy as np
import pytest

from oracle import pauli_oracle as orc

H2 = {'paulis': [{"coeff": {"imag": 0.0, "real": -1.052373245772859}, "label": "II"},
                 {"coeff": {"imag": 0.0, "real": 0.39793742484318045}, "label": "ZI"},
                 {"coeff": {"imag": 0.0, "real": -0.39793742484318045}, "label": "IZ"},
                 {"coeff": {"imag": 0.0, "real": -0.01128010425623538}, "label": "ZZ"},
                 {"coeff": {"imag": 0.0, "real": 0.18093119978423156}, "label": "XX"}]}
AUX1 = {'paulis': [{'coeff': {'imag': 0.0, 'real': 2.0}, 'label': 'II'}]}
AUX2 = {'paulis': [{'coeff': {'imag': 0.0, 'real': 0.5}, 'label': 'II'},
                   {'coeff': {'imag': 0.0, 'real': 0.5}, 'label': 'ZZ'},
                   {'coeff': {'imag': 0.0, 'real': 0.5}, 'label': 'YY'},
                   {'coeff': {'imag': 0.0, 'real': -0.5}, 'label': 'XX'}]}
G1_SPECTRUM = [-1.85727503, -1.24458455, -0.88272215, -0.22491125]


def test_g1_h2_spectrum():
    # test/aqua/test_numpy_eigen_solver.py:26-53
    terms = orc.from_dict(H2)
    val, vec = orc.ref_solve(terms, k=1)
    assert abs(val[0] - (-1.85727503)) < 1e-7
    val, vec = orc.ref_solve(terms, k=4)
    np.testing.assert_array_almost_equal(val.real, G1_SPECTRUM)
    assert vec.shape == (4, 4)


def test_g2_aux_ops_on_ground_state():
    # test/aqua/test_numpy_minimum_eigen_solver.py:39-58
    terms = orc.from_dict(H2)
    val, vec = orc.ref_solve(terms, k=1)
    a1 = orc.ref_eval_aux(orc.from_dict(AUX1), vec[0])
    a2 = orc.ref_eval_aux(orc.from_dict(AUX2), vec[0])
    np.testing.assert_array_almost_equal(a1, [2, 0])
    np.testing.assert_array_almost_equal(a2, [0, 0])


def test_g3_filter():
    # test_numpy_eigen_solver.py:61-74 (v >= -1 keeps the top two); dense branch, all 4 values
    terms = orc.from_dict(H2)
    val, _ = orc.ref_solve(terms, k=4)
    kept = [v.real for v in val if v.real >= -1]
    np.testing.assert_array_almost_equal(kept, [-0.88272215, -0.22491125])
    kept = [v.real for v in val if v.real >= -0.5]
    np.testing.assert_array_almost_equal(kept[:1], [-0.22491125])


def test_g4_single_pauli_elements_and_ordering():
    # test/aqua/operators/test_op_construction.py:72-83, 95-106, 158-162
    X = orc.pauli_csr('X').toarray()
    Y = orc.pauli_csr('Y').toarray()
    Z = orc.pauli_csr('Z').toarray()
    assert Z[0, 0] == 1 and Z[1, 1] == -1
    assert X[0, 1] == 1 and X[1, 0] == 1
    assert Y[0, 1] == -1j          # Y.eval('1').eval('0') == -1j  (<0|Y|1>)
    assert Y[1, 0] == 1j
    np.testing.assert_array_equal(orc.pauli_csr('XY').toarray(), np.kron(X, Y))
    np.testing.assert_array_equal(orc.pauli_csr('ZIXY').toarray(),
                                  np.kron(np.kron(np.kron(Z, np.eye(2)), X), Y))


def test_g5_expectation_table():
    # test/aqua/operators/expectations/test_matrix_expectation.py:121-134
    s = 1 / np.sqrt(2)
    states = {'plus': np.array([s, s]), 'minus': np.array([s, -s]),
              'zero': np.array([1, 0]), 'one': np.array([0, 1])}
    valids = {'plus': [1, 0, 0, 1], 'minus': [-1, 0, 0, 1], 'zero': [0, 0, 1, 1],
              'one': [0, 0, -1, 1]}
    for name, psi in states.items():
        got = [orc.ref_matrix_expectation([(1.0, p)], psi.astype(complex)) for p in 'XYZI']
        np.testing.assert_array_almost_equal(np.real(got), valids[name], decimal=12)


def test_g6_state_overlaps():
    # test/aqua/operators/test_state_op_meas_evals.py:32-50
    n = 6
    psi = np.zeros(1 << n, dtype=complex)
    psi[int('101010', 2)] += 4 * .5
    psi[int('111111', 2)] += 4 * .3
    psi[0] += 3 + .1j
    assert abs(np.vdot(psi, psi) - 14.45) < 1e-12
    ghz = np.zeros(16, dtype=complex)
    ghz[0] = ghz[15] = 1 / np.sqrt(2)
    assert abs(orc.ref_evaluate_with_statevector([(1.0, 'XXXX')], ghz)[0] - 1) < 1e-12


def test_g7_maxcut():
    # test/optimization/test_qaoa.py:33-52 (W2) and test_classical_cplex.py:28-44
    w2 = np.array([[0., 8., -9., 0.], [8., 0., 7., 9.], [-9., 7., 0., -8.], [0., 9., -8., 0.]])
    terms, _ = orc.max_cut_operator(w2)
    val, vec = orc.ref_solve(terms, k=1)
    x = orc.sample_most_likely(vec[0])
    sol = ''.join(str(int(1 - b)) for b in x)           # get_graph_solution: 1 - x
    assert sol in {'1011', '0100'} or ''.join(str(int(b)) for b in x) in {'1011', '0100'}
    w = orc.random_graph(4, edge_prob=0.5, weight_range=10, seed=8123179999)
    terms, offset = orc.max_cut_operator(w)
    val, vec = orc.ref_solve(terms, k=1)
    assert abs(val[0].real - (-9.0)) < 1e-12            # energy -9.0
    x = orc.sample_most_likely(vec[0])
    cut = sum(w[i, j] * x[i] * (1 - x[j]) for i in range(4) for j in range(4))
    assert cut == 10                                     # max_cut_value 10


def test_g9_column_form_matches_row_form():
    # primitive_ops/pauli_op.py:204-218: out[b ^ x] += v * prod(1-2(b&z)) * (+i)^{nY};
    # checked against the dense matrix for Z^I^X^Y (test_op_construction.py:108-116)
    lab = 'ZIXY'
    x, z = orc.label_to_masks(lab)
    ny = bin(x & z).count('1')
    dense = orc.pauli_csr(lab).toarray()
    for b in range(16):
        col = np.zeros(16, dtype=complex)
        col[b ^ x] = (1j) ** ny * (-1) ** bin(b & z).count('1')
        np.testing.assert_array_almost_equal(dense[:, b], col)


def test_g12_pauli_product_phase():
    # test_weighted_pauli_operator.py:98-112: 0.5 IXYZ * 0.5 ZYIX = -0.25 ZZYY
    (cx, cz), ph = orc.pauli_dot(orc.label_to_masks('IXYZ'), orc.label_to_masks('ZYIX'))
    assert orc.masks_to_label(cx, cz, 4) == 'ZZYY'
    assert 0.25 * ph == -0.25
    # brute force all 2-qubit pairs against dense products
    labs = [''.join(p) for p in itertools.product('IXYZ', repeat=2)]
    for a in labs:
        for b in labs:
            (cx, cz), ph = orc.pauli_dot(orc.label_to_masks(a), orc.label_to_masks(b))
            lhs = orc.pauli_csr(a).toarray() @ orc.pauli_csr(b).toarray()
            np.testing.assert_array_almost_equal(lhs, ph * orc.pauli_csr(
                orc.masks_to_label(cx, cz, 2)).toarray())


def test_g13_csr_equals_matrix_free_and_per_pauli_sum():
    # test_weighted_pauli_operator.py:35-44, 483-496: seed-3 random 3-qubit 64-term operator;
    # evaluate_with_statevector == sum_k w_k vdot(psi, P_k psi) to 10 places
    rng = np.random.default_rng(3)
    labs = [''.join(p) for p in itertools.product('IXYZ', repeat=3)]
    weights = rng.random(len(labs))
    terms = orc.simplify(list(zip(weights, labs)))
    psi = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    psi /= np.linalg.norm(psi)
    ref = orc.ref_evaluate_with_statevector(terms, psi)[0]
    per_pauli = sum(w * np.vdot(psi, orc.pauli_csr(lab).dot(psi)) for w, lab in terms)
    assert abs(ref - per_pauli) < 1e-10
    xm, zm, yp, cf = orc.pack_terms(terms)
    assert abs(ref - orc.mf_expect(3, xm, zm, yp, cf, psi)) < 1e-12
    np.testing.assert_allclose(orc.mf_apply(3, xm, zm, yp, cf, psi),
                               orc.ref_to_spmatrix(terms).dot(psi), atol=1e-13)


def _chem(golden_dir, key):
    with open(os.path.join(golden_dir, 'chemistry_integrals.json')) as f:
        return json.load(f)[key]


def test_g10_h2_fcidump_chain(golden_dir):
    g = _chem(golden_dir, 'G10_h2_fcidump')
    h1 = orc.onee_to_spin(np.array(g['mo_onee']))
    h2 = orc.twoe_to_spin(np.array(g['mo_eri']))
    terms = orc.jordan_wigner(h1, h2)
    assert len(terms[0][1]) == g['expect_num_qubits']
    assert len(terms) == g['expect_num_paulis']
    ev = np.linalg.eigvalsh(orc.ref_to_spmatrix(terms).toarray())
    assert abs(ev[0] - g['expect_min_eigenvalue']) < 1e-7
    for v in G1_SPECTRUM:                                  # 2-qubit tapered spectrum is contained
        assert np.min(np.abs(ev - v)) < 1e-6


@pytest.mark.timeout(600)
def test_g11_lih_chain(golden_dir):
    g = _chem(golden_dir, 'G11_lih_npz')
    h1 = orc.onee_to_spin(np.array(g['mo_onee']))
    h2 = orc.twoe_to_spin(np.array(g['mo_eri']))
    terms = orc.jordan_wigner(h1, h2)
    assert len(terms[0][1]) == 12
    xm, zm, yp, cf = orc.pack_terms(terms)
    assert np.max(np.abs(cf.imag)) < 1e-12
    import scipy.sparse.linalg as spla
    N = 1 << 12
    op = spla.LinearOperator((N, N), dtype=np.complex128,
                             matvec=lambda v: orc.mf_apply(12, xm, zm, yp, cf, v.astype(complex)))
    e0 = spla.eigsh(op, k=1, which='SA', return_eigenvectors=False)[0]
    assert abs(round(e0 + g['nuclear_repulsion'], 3) - g['expect_total_energy_3dp']) < 1.5e-3
    assert len(set(xm.tolist())) < len(xm) / 3            # few distinct x-masks (grouping pays)
