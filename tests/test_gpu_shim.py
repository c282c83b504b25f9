"""GPU: the reference-facing Python entry points (drop-in boundary) against the reference's own
known answers and the oracle.  These read like the reference's tests:
test_weighted_pauli_operator.py, test_numpy_eigen_solver.py, test_numpy_minimum_eigen_solver.py,
test_matrix_expectation.py, test_state_op_meas_evals.py, test_qaoa.py / test_classical_cplex.py."""
import itertools

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import pauli_oracle as po        # noqa: E402
from qiskit_aqua_b200 import (AquaError, I, ListOp, MatrixExpectation, NumPyEigensolver, NumPyMinimumEigensolver,  # noqa: E402
                              Pauli, StateFn, WeightedPauliOperator, X, Y, Z, max_cut, random_graph,
                              sample_most_likely)

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


class _StatevectorResult:
    """Stand-in for a backend Result holding the statevectors of construct_evaluation_circuit
    (statevector mode): 'psi' and one circuit per non-identity Pauli, named by its label."""

    def __init__(self, psi, terms, prefix=''):
        self._sv = {prefix + 'psi': psi}
        for _, lab in terms:
            if set(lab) != {'I'}:
                self._sv[prefix + lab] = po.pauli_csr(lab).dot(psi)

    def get_statevector(self, name):
        return self._sv[name]

    def data(self, name):
        return {'snapshots': {'expectation_value': {'expval': [{'value': [0.25, -0.5]}]}}}


def test_evaluate_with_result_statevector_mode():
    # mirrors test_weighted_pauli_operator.py:483-496 (test_evaluate_statevector_mode): the value from
    # the per-Pauli statevectors equals evaluate_with_statevector to 10 places
    rng = np.random.default_rng(3)
    paulis = [Pauli(''.join(lab)) for lab in itertools.product('IXYZ', repeat=3)]
    weights = rng.random(len(paulis)) + 1j * rng.random(len(paulis))
    op = WeightedPauliOperator.from_list(paulis, weights)
    psi = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    psi /= np.linalg.norm(psi)
    terms = [(w, p.to_label()) for w, p in op.paulis]
    reference = op.evaluate_with_statevector(psi)
    for prefix in ('', 'pre_'):
        actual = op.evaluate_with_result(_StatevectorResult(psi, terms, prefix), statevector_mode=True,
                                         circuit_name_prefix=prefix)
        assert abs(reference[0] - actual[0]) < 1e-10 and actual[1] == 0.0
    snap = op.evaluate_with_result(_StatevectorResult(psi, terms), statevector_mode=True, use_simulator_snapshot_mode=True)
    assert snap[0] == 0.25 - 0.5j
    with pytest.raises(AquaError):
        op.evaluate_with_result(_StatevectorResult(psi, terms), statevector_mode=False)
    with pytest.raises(AquaError):
        WeightedPauliOperator(paulis=[]).evaluate_with_result(None, True)


def test_evaluate_with_statevector_g13():
    # seed-3 random 3-qubit 64-term operator (test_weighted_pauli_operator.py:35-44, 483-496)
    rng = np.random.default_rng(3)
    paulis = [Pauli(''.join(lab)) for lab in itertools.product('IXYZ', repeat=3)]
    weights = rng.random(len(paulis))
    op = WeightedPauliOperator.from_list(paulis, weights)
    psi = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    psi /= np.linalg.norm(psi)
    mean, std = op.evaluate_with_statevector(psi)
    terms = [(w, p.to_label()) for w, p in op.paulis]
    ref = po.ref_evaluate_with_statevector(terms, psi)[0]
    assert std == 0.0
    assert abs(mean - ref) < 1e-10                      # places=10 in the reference test
    per_pauli = sum(w * np.vdot(psi, po.pauli_csr(lab).dot(psi)) for w, lab in terms)
    assert abs(mean - per_pauli) < 1e-10


def test_evaluate_with_statevector_complex_weights_16q():
    rng = np.random.default_rng(5)
    n = 16
    labels = set()
    while len(labels) < 120:
        labels.add(''.join(rng.choice(list('IXYZ'), p=[.7, .1, .1, .1]) for _ in range(n)))
    labels = sorted(labels)
    weights = rng.standard_normal(120) + 1j * rng.standard_normal(120)
    op = WeightedPauliOperator.from_list([Pauli(l) for l in labels], list(weights))
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    x, z, y, c = po.pack_terms(list(zip(weights, labels)))
    ref = po.mf_expect(n, x, z, y, c, psi)
    got = op.evaluate_with_statevector(torch.from_numpy(psi).cuda())[0]     # zero-copy device input
    assert abs(got - ref) <= 1e-10 * abs(ref)


def test_numpy_eigensolver_g1_g2_g3():
    qubit_op = WeightedPauliOperator.from_dict(H2)
    res = NumPyEigensolver(qubit_op, aux_operators=[]).run()
    assert len(res.eigenvalues) == 1 and len(res.eigenstates) == 1
    assert abs(res.eigenvalues[0] - (-1.85727503 + 0j)) < 1e-7
    res = NumPyEigensolver(qubit_op, k=4, aux_operators=[]).run()
    assert len(res.eigenvalues) == 4 and len(res.eigenstates) == 4
    np.testing.assert_array_almost_equal(res.eigenvalues.real,
                                         [-1.85727503, -1.24458455, -0.88272215, -0.22491125])
    # filter v >= -1 keeps the upper two (test_numpy_eigen_solver.py:61-74)
    res = NumPyEigensolver(qubit_op, k=2, aux_operators=[], filter_criterion=lambda x, v, a: v >= -1).run()
    np.testing.assert_array_almost_equal(res.eigenvalues.real, [-0.88272215, -0.22491125])
    res = NumPyEigensolver(qubit_op, k=1, filter_criterion=lambda x, v, a: False).run()
    assert len(res.eigenvalues) == 0


def test_numpy_minimum_eigensolver_aux_ops():
    qubit_op = WeightedPauliOperator.from_dict(H2)
    aux = [WeightedPauliOperator.from_dict(AUX1), WeightedPauliOperator.from_dict(AUX2)]
    res = NumPyMinimumEigensolver(qubit_op, aux_operators=aux).run()
    assert abs(res.eigenvalue - (-1.85727503 + 0j)) < 1e-7
    assert len(res.aux_operator_eigenvalues) == 2
    np.testing.assert_array_almost_equal(res.aux_operator_eigenvalues[0], [2, 0])
    np.testing.assert_array_almost_equal(res.aux_operator_eigenvalues[1], [0, 0])
    # reuse with a new operator (test_numpy_minimum_eigen_solver.py:66-109)
    algo = NumPyMinimumEigensolver()
    res = algo.compute_minimum_eigenvalue(qubit_op)
    assert abs(res.eigenvalue - (-1.85727503 + 0j)) < 1e-7 and res.aux_operator_eigenvalues is None
    res = algo.compute_minimum_eigenvalue(qubit_op, aux_operators=aux)
    np.testing.assert_array_almost_equal(res.aux_operator_eigenvalues[0], [2, 0])
    res = NumPyMinimumEigensolver(qubit_op, filter_criterion=lambda x, v, a: v >= -0.5).run()
    assert abs(res.eigenvalue - (-0.22491125 + 0j)) < 1e-7


def test_matrix_expectation_table_g5():
    s = 1 / np.sqrt(2)
    states = {'plus': [s, s], 'minus': [s, -s], 'zero': [1, 0], 'one': [0, 1]}
    valids = {'plus': [1, 0, 0, 1], 'minus': [-1, 0, 0, 1], 'zero': [0, 0, 1, 1], 'one': [0, 0, -1, 1]}
    expect = MatrixExpectation()
    for name, vec in states.items():
        psi = StateFn(np.array(vec, dtype=complex))
        got = [expect.convert(~StateFn(op) @ psi).eval() for op in (X, Y, Z, I)]
        np.testing.assert_array_almost_equal(np.real(got), valids[name], decimal=12)
    # ListOp of states against one measurement (matrix_op.py:199-201 broadcast)
    meas = ~StateFn(Z)
    vals = meas.eval(ListOp([StateFn(np.array(v, dtype=complex)) for v in states.values()]))
    np.testing.assert_array_almost_equal(np.real(vals), [0, 0, 1, -1], decimal=12)


def test_state_overlaps_g6():
    n = 6
    psi = np.zeros(1 << n, dtype=complex)
    psi[int('101010', 2)] += 4 * .5
    psi[int('111111', 2)] += 4 * .3
    psi[0] += 3 + .1j
    sf = StateFn(psi)
    assert abs((~sf).eval(sf) - 14.45) < 1e-12
    # the reference builds it from dict states: 4*{'101010': .5, '111111': .3} + (3+.1j)*Zero^6
    sf2 = 4 * StateFn({'101010': .5, '111111': .3}) + (3 + .1j) * StateFn('000000')
    assert abs((~sf2).eval(sf2) - 14.45) < 1e-12
    ghz = np.zeros(16, dtype=complex)
    ghz[0] = ghz[15] = 1 / np.sqrt(2)
    assert abs((~StateFn(X ^ 4) @ StateFn(ghz)).eval() - 1) < 1e-12
    y1 = Y.eval(StateFn(np.array([0, 1], dtype=complex))).to_matrix()      # Y|1> = -i|0>
    np.testing.assert_array_almost_equal(y1, [-1j, 0])


def test_maxcut_g7_diagonal_branch():
    w2 = np.array([[0., 8., -9., 0.], [8., 0., 7., 9.], [-9., 7., 0., -8.], [0., 9., -8., 0.]])
    op, _ = max_cut.get_operator(w2)
    res = NumPyMinimumEigensolver(op).run()
    x = sample_most_likely(res.eigenstate)
    sol = ''.join(str(int(b)) for b in max_cut.get_graph_solution(x))
    assert sol in {'1011', '0100'}
    w = random_graph(4, edge_prob=0.5, weight_range=10, seed=8123179999)
    op, offset = max_cut.get_operator(w)
    res = NumPyMinimumEigensolver(op).run()
    assert abs(res.eigenvalue.real - (-9.0)) < 1e-12
    x = sample_most_likely(res.eigenstate)
    assert max_cut.max_cut_value(x, w) == 10
    # config 1 size: 16 nodes, k = 3, against the oracle's sort branch
    w = random_graph(16, weight_range=10, edge_prob=0.5, negative_weight=True, seed=17)
    op, _ = max_cut.get_operator(w)
    res = NumPyEigensolver(op, k=3).run()
    terms, _ = po.max_cut_operator(w)
    ref, _ = po.ref_solve(terms, k=3)
    np.testing.assert_allclose(np.real(res.eigenvalues), ref.real, atol=1e-12)


def test_to_matrix_and_massive_guard():
    # test_op_construction.py:158-162 ((X^Y).to_matrix() == kron(X, Y)) and operator_base.py:634-648
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    y = np.array([[0, -1j], [1j, 0]], dtype=complex)
    np.testing.assert_allclose((X ^ Y).to_matrix(), np.kron(x, y), atol=1e-15)
    h = 0.5 * (X ^ X ^ I) + (Z ^ I ^ Y) - 0.25j * (Y ^ Y ^ Z)
    terms = [(w, p.to_label()) for w, p in h.to_legacy_op().paulis]
    np.testing.assert_allclose(h.to_matrix(), po.ref_to_spmatrix(terms).toarray(), atol=1e-14)
    np.testing.assert_allclose(h.to_spmatrix().toarray(), po.ref_to_spmatrix(terms).toarray(), atol=1e-14)
    with pytest.raises(ValueError):
        (X ^ 17).to_matrix()
