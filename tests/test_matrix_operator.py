"""Legacy MatrixOperator / op_converter (qiskit_aqua_b200/legacy.py) against
outputs of the reference's own classes (tests/golden/reference_legacy.json, made by
tests/golden/make_legacy_fixtures.py) and the reference's round-trip test
(test/aqua/operators/legacy/test_op_converter.py:30-58)."""
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q
from qiskit_aqua_b200.legacy import MatrixOperator, op_converter, pauli_decompose
from oracle import pauli_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, 'golden', 'reference_legacy.json')))


def mat(rec):
    return np.array(rec['re']) + 1j * np.array(rec['im'])


def same_paulis(got_paulis, want, tol=1e-12):
    got = [(p.to_label(), complex(w)) for w, p in got_paulis]
    want = [(lab, complex(*w)) for w, lab in want]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=tol, atol=tol)


# ---- matrix -> Pauli sum: one Walsh-Hadamard transform per x-mask on the device ---------------------
@pytest.mark.gpu
@pytest.mark.parametrize('rec', FX['decompose'], ids=[r['name'] for r in FX['decompose']])
def test_matrix_to_pauli_sum_equals_reference(rec):
    """tr(P M)/2^n by the device Walsh-Hadamard kernel == the reference's 4^n sparse products, same order."""
    mo = MatrixOperator(mat(rec['matrix']))
    assert mo.num_qubits == rec['num_qubits'] and str(mo) == rec['str']
    same_paulis(op_converter.to_weighted_pauli_operator(mo).paulis, rec['paulis'])
    # and the decomposition really is the matrix (independent of the fixtures)
    terms = [(complex(w), p.to_label()) for w, p in mo.pauli_sum.paulis]
    assert np.allclose(po.ref_to_spmatrix(terms).toarray(), mat(rec['matrix']), atol=1e-12)


@pytest.mark.gpu
def test_decompose_options_and_guards():
    m = np.diag([1.0, 2.0, 3.0, 5.0])
    full = pauli_decompose(m)
    diag_only = pauli_decompose(m, diagonal_only=True)
    assert [(p.to_label(), w) for w, p in full] == [(p.to_label(), w) for w, p in diag_only]
    assert [p.to_label() for _, p in full] == ['II', 'IZ', 'ZI', 'ZZ']
    assert pauli_decompose(np.zeros((2, 2))) == []
    with pytest.raises(q.AquaError):
        pauli_decompose(np.ones((3, 3)))
    with pytest.raises(ValueError):
        pauli_decompose(np.zeros((1 << 13, 1 << 13), dtype=np.int8))
    assert MatrixOperator(None).is_empty() and MatrixOperator(np.zeros((4, 4))).is_empty()
    assert MatrixOperator(None).num_qubits == 0
    with pytest.raises(q.AquaError):
        MatrixOperator(None).evaluate_with_statevector(np.ones(2))


# ---- host side (no device) ----------------------------------------------------------------------
def test_converters_do_not_build_a_matrix():
    rec = FX['pauli_sums'][0]
    op = q.WeightedPauliOperator(paulis=[[complex(*w), q.Pauli(l)] for w, l in rec['terms']])
    mo = op_converter.to_matrix_operator(op)
    assert mo.pauli_sum is op and mo.num_qubits == 2        # shared, nothing assembled
    assert op_converter.to_matrix_operator(mo) is mo
    assert op_converter.to_weighted_pauli_operator(op) is op
    same_paulis(op_converter.to_weighted_pauli_operator(mo).paulis, rec['back'])
    assert op_converter.to_matrix_operator(q.WeightedPauliOperator(paulis=[])).is_empty()
    with pytest.raises(q.AquaError):
        op_converter.to_matrix_operator('nope')
    # algebra on the Pauli sums is the matrix algebra
    mo2 = op_converter.to_matrix_operator(q.WeightedPauliOperator(paulis=[[complex(*w), q.Pauli(l)] for w, l in rec['other']]))
    for got, key in (((mo + mo2), 'add'), ((mo - mo2), 'sub'), ((mo * mo2), 'mul'), ((-mo), 'neg')):
        terms = [(complex(w), p.to_label()) for w, p in got.pauli_sum.paulis]
        assert np.allclose(po.ref_to_spmatrix(terms).toarray(), mat(rec[key]), atol=1e-12), key
    assert mo == mo.copy() and not (mo == mo2)


@pytest.mark.parametrize('rec', FX['pauli_sums'], ids=[r['name'] for r in FX['pauli_sums']])
def test_pauli_sum_round_trip_order(rec):
    op = q.WeightedPauliOperator(paulis=[[complex(*w), q.Pauli(l)] for w, l in rec['terms']])
    same_paulis(op_converter.to_weighted_pauli_operator(op_converter.to_matrix_operator(op)).paulis, rec['back'])


# ---- device ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('rec', FX['decompose'], ids=[r['name'] for r in FX['decompose']])
def test_matrix_operator_on_device_equals_reference(rec):
    m = mat(rec['matrix'])
    mo = MatrixOperator(m)
    psi = mat(rec['psi']).reshape(-1)
    mean, std = mo.evaluate_with_statevector(psi)
    assert std == 0.0
    assert mean == pytest.approx(complex(*rec['expectation']), rel=1e-10, abs=1e-12)
    assert np.allclose(mo.dense_matrix, m, atol=1e-12)
    assert (mo.dia_matrix is not None) == rec['is_diagonal']
    want = mat(rec['matrix_property'])
    got = mo.matrix
    if rec['is_diagonal']:
        assert np.allclose(np.asarray(got).reshape(1, -1), want, atol=1e-12)
    else:
        assert np.allclose(got.toarray(), want, atol=1e-12)
    chopped = mo.chop(0.3, copy=True)
    assert chopped.is_empty() == rec['chop_0.3_empty']
    if not chopped.is_empty():
        assert np.allclose(chopped.dense_matrix, mat(rec['chop_0.3']), atol=1e-12)
    assert np.allclose(mo.dense_matrix, m, atol=1e-12)           # copy=True left the original alone


@pytest.mark.gpu
@pytest.mark.parametrize('rec', FX['pauli_sums'], ids=[r['name'] for r in FX['pauli_sums']])
def test_to_matrix_operator_views_on_device(rec):
    op = q.WeightedPauliOperator(paulis=[[complex(*w), q.Pauli(l)] for w, l in rec['terms']])
    mo = op_converter.to_matrix_operator(op)
    assert np.allclose(mo.dense_matrix, mat(rec['dense']), atol=1e-12)
    assert (mo.dia_matrix is not None) == rec['is_diagonal']


@pytest.mark.gpu
def test_reference_round_trip_test_on_device():
    """test_op_converter.py:30-58: random 2-qubit 16-term Pauli sum -> matrix -> Pauli sum, 8 decimals."""
    rng = np.random.default_rng(0)
    labels = [''.join(t) for t in __import__('itertools').product('IXYZ', repeat=2)]
    weights = rng.random(16)
    op = q.WeightedPauliOperator.from_list([q.Pauli(l) for l in labels], weights)
    mo = op_converter.to_matrix_operator(op)
    back = op_converter.to_weighted_pauli_operator(MatrixOperator(mo.matrix))
    got = {p.to_label(): complex(w) for w, p in back.paulis}
    for lab, w in zip(labels, weights):
        assert got[lab] == pytest.approx(w, abs=1e-8)
    # and the exact evolution branch agrees with expm of the same matrix
    import scipy.linalg as sla
    psi = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    psi /= np.linalg.norm(psi)
    assert np.allclose(mo.evolve(psi, 0.7, 0), sla.expm(-0.7j * mo.dense_matrix) @ psi, atol=1e-8)
