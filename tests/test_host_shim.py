"""CPU-only: the host-side mirror of the reference interface (container semantics, errors, input
generators).  Mirrors reference tests test_weighted_pauli_operator.py / test_op_construction.py."""
import json
import os

import numpy as np
import pytest

from oracle import pauli_oracle as po
from qiskit_aqua_b200 import (AquaError, FermionicOperator, I, NumPyEigensolver, Pauli, PauliOp, SummedOp,
                              WeightedPauliOperator, X, Y, Z, max_cut, random_graph, sample_most_likely)


def test_pauli_conventions():
    p = Pauli('IXZY')
    assert p.num_qubits == 4 and p.to_label() == 'IXZY'
    np.testing.assert_array_equal(p.x, [True, False, True, False])      # little-endian arrays
    np.testing.assert_array_equal(p.z, [True, True, False, False])
    q = Pauli((p.z, p.x))
    assert q == p and hash(q) == hash(p)
    assert (p.x_mask, p.z_mask) == po.label_to_masks('IXZY')
    with pytest.raises(ValueError):
        Pauli('IXA')


def test_pauli_product_phase_g12():
    # test_weighted_pauli_operator.py:98-112
    a = WeightedPauliOperator(paulis=[[0.5, Pauli('IXYZ')]])
    b = WeightedPauliOperator(paulis=[[0.5, Pauli('ZYIX')]])
    c = a * b
    assert len(c.paulis) == 1
    assert c.paulis[0][1].to_label() == 'ZZYY' and c.paulis[0][0] == -0.25


def test_simplify_chop_and_dict_roundtrip(tmp_path):
    op = WeightedPauliOperator.from_list([Pauli('XX'), Pauli('ZZ'), Pauli('XX'), Pauli('YY')],
                                         [0.5, 1e-14, -0.5, 0.25j])
    labels = [p.to_label() for _, p in op.paulis]
    assert labels == ['ZZ', 'YY']                                     # XX cancelled exactly
    op.chop(1e-12)
    assert [p.to_label() for _, p in op.paulis] == ['YY']
    f = tmp_path / 'op.json'
    op.to_file(str(f))
    op2 = WeightedPauliOperator.from_file(str(f))
    assert op2 == op and op2.num_qubits == 2
    with pytest.raises(AquaError):
        WeightedPauliOperator.from_dict({'nope': []})
    with pytest.raises(ValueError):
        WeightedPauliOperator.from_list([Pauli('X')], [1.0, 2.0])


def test_empty_operator_errors():
    op = WeightedPauliOperator(paulis=[])
    assert op.is_empty()
    with pytest.raises(AquaError):
        op.evaluate_with_statevector(np.ones(2, dtype=complex))
    with pytest.raises(AquaError):
        NumPyEigensolver().run()
    with pytest.raises(ValueError):
        NumPyEigensolver(k=0)


def test_opflow_algebra():
    h = 0.5 * (X ^ X) + (Z ^ I) - 0.25 * (Y ^ Y)
    assert isinstance(h, SummedOp) and h.num_qubits == 2
    terms = {p.to_label(): w for w, p in h.to_legacy_op().paulis}
    assert terms == {'XX': 0.5, 'ZI': 1.0, 'YY': -0.25}
    assert (X ^ 3).primitive.to_label() == 'XXX'
    assert isinstance(X ^ Y, PauliOp) and (X ^ Y).primitive.to_label() == 'XY'
    legacy = WeightedPauliOperator(paulis=[[1.0, Pauli('XZ')], [2.0, Pauli('ZZ')]])
    flow = legacy.to_opflow()
    assert {o.primitive.to_label(): o.coeff for o in flow.oplist} == {'XZ': 1.0, 'ZZ': 2.0}


def test_maxcut_generator_matches_oracle():
    w = random_graph(6, edge_prob=0.5, weight_range=10, seed=99)
    np.testing.assert_array_equal(w, po.random_graph(6, edge_prob=0.5, weight_range=10, seed=99))
    op, shift = max_cut.get_operator(w)
    ref_terms, ref_shift = po.max_cut_operator(w)
    assert shift == ref_shift
    assert [(complex(a), p.to_label()) for a, p in op.paulis] == [(complex(a), l) for a, l in ref_terms]
    v = np.zeros(8)
    v[5] = 1.0
    np.testing.assert_array_equal(sample_most_likely(v), [1, 0, 1])


def test_jordan_wigner_generator_matches_oracle(golden_dir):
    g = json.load(open(os.path.join(golden_dir, 'chemistry_integrals.json')))['G10_h2_fcidump']
    h1 = po.onee_to_spin(np.array(g['mo_onee']))
    h2 = po.twoe_to_spin(np.array(g['mo_eri']))
    op = FermionicOperator(h1, h2).mapping('jordan_wigner')
    ref = po.jordan_wigner(h1, h2)
    assert len(op.paulis) == 15
    for (w, p), (rw, rl) in zip(op.paulis, ref):
        assert p.to_label() == rl and abs(w - rw) < 1e-15
    rng = np.random.default_rng(14)
    h1 = rng.standard_normal((5, 5))
    h1 = h1 + h1.T
    h2 = np.zeros((5,) * 4)
    for _ in range(12):
        i, j, k, m = rng.integers(0, 5, size=4)
        h2[i, j, k, m] = rng.standard_normal()
    op = FermionicOperator(h1, h2).mapping('jordan_wigner')
    ref = po.jordan_wigner(h1, h2)
    assert [(p.to_label()) for _, p in op.paulis] == [l for _, l in ref]
    np.testing.assert_allclose([w for w, _ in op.paulis], [w for w, _ in ref], atol=1e-15)


# ---------------------------------------------------------------------------------------------
# the reference's own container tests (test/aqua/operators/legacy/test_weighted_pauli_operator.py)
# ---------------------------------------------------------------------------------------------
SIX = ['IXYZ', 'XXZY', 'IIZZ', 'XXYY', 'ZZXX', 'YYYY']


def _term(label, w):
    return WeightedPauliOperator(paulis=[[w, Pauli(label)]])


def test_ref_from_to_file_num_qubits_is_empty_str(tmp_path):
    weights = [0.2 + -1j * 0.8, 0.6 + -1j * 0.6, 0.8 + -1j * 0.2, -0.2 + -1j * 0.8, -0.6 - -1j * 0.6, -0.8 - -1j * 0.2]
    op = WeightedPauliOperator.from_list([Pauli(x) for x in SIX], weights)
    path = str(tmp_path / 'temp_op.json')
    op.to_file(path)                                         # :61-73
    assert WeightedPauliOperator.from_file(path) == op
    empty = WeightedPauliOperator(paulis=[])
    assert empty.num_qubits == 0 and empty.is_empty()        # :75-85
    assert op.num_qubits == 4 and not op.is_empty()
    op_a = _term('IXYZ', 0.5)
    op_a += _term('ZYIX', 0.5)
    assert str(op_a) == 'Representation: paulis, qubits: 4, size: 2'          # :87-101
    assert str(WeightedPauliOperator(paulis=[[0.5, Pauli('IXYZ')]], name='ABC')) == \
        'ABC: Representation: paulis, qubits: 4, size: 1'


def test_ref_multiplication():
    new_op = _term('IXYZ', 0.5) * _term('ZYIX', 0.5)         # :103-122
    assert len(new_op.paulis) == 1
    assert new_op.paulis[0][0] == -0.25 and new_op.paulis[0][1].to_label() == 'ZZYY'
    new_op = -2j * new_op
    assert new_op.paulis[0][0] == 0.5j
    new_op = new_op * 0.3j
    assert new_op.paulis[0][0] == -0.15


def test_ref_add_sub_in_place_and_copy():
    op_a, op_b = _term('IXYZ', 0.5), _term('ZYIX', 0.5)
    ori_a, ori_b = op_a.copy(), op_b.copy()
    op_a += op_b                                             # :124-146
    assert op_a != ori_a and op_b == ori_b and len(op_a.paulis) == 2
    op_a += _term('IXYZ', 0.25)
    assert len(op_a.paulis) == 2 and op_a.paulis[0][0] == 0.75
    op_a, op_b = _term('IXYZ', 0.5), _term('ZYIX', 0.5)
    new_op = op_a + op_b                                     # :148-171
    assert op_a == ori_a and op_b == ori_b and len(op_a.paulis) == 1 and len(new_op.paulis) == 2
    new_op = new_op + _term('IXYZ', 0.25)
    assert len(new_op.paulis) == 2 and new_op.paulis[0][0] == 0.75
    new_op = op_a - op_b                                     # :173-198
    assert op_a == ori_a and op_b == ori_b and len(new_op.paulis) == 2
    assert new_op.paulis[0][0] == 0.5 and new_op.paulis[1][0] == -0.5
    new_op = new_op - _term('IXYZ', 0.25)
    assert len(new_op.paulis) == 2 and new_op.paulis[0][0] == 0.25
    op_a -= op_b                                             # :200-222
    assert op_a != ori_a and op_b == ori_b and len(op_a.paulis) == 2
    op_a = _term('IXYZ', 0.5)
    op_a -= _term('ZYIX', 0.5)
    op_a -= _term('IXYZ', 0.5)
    assert len(op_a.paulis) == 2                             # "sub does not remove zero weights"
    with pytest.raises(AquaError):
        _term('IXYZ', 1.0) + _term('XX', 1.0)


def test_ref_equal_negation_simplify():
    paulis = [Pauli(x) for x in SIX]
    coeffs = [0.2, 0.6, 0.8, -0.2, -0.6, -0.8]
    op1 = WeightedPauliOperator.from_list(paulis, coeffs)
    op2 = WeightedPauliOperator.from_list(paulis, list(coeffs))
    op3 = WeightedPauliOperator.from_list(paulis, [-c for c in coeffs])
    op4 = WeightedPauliOperator.from_list(paulis, [-0.2, 0.6, 0.8, -0.2, -0.6, -0.8])
    assert op1 == op2 and op1 != op3 and op1 != op4 and op3 != op4            # :224-240
    assert op1 == -op3 and -op1 == op3                                          # :242-256
    assert op1 * -1.0 == op3
    new_op = _term('IXYZ', 0.5) + _term('IXYZ', -0.5)                           # :258-283
    new_op.simplify()
    assert len(new_op.paulis) == 0 and new_op.is_empty()
    for i, pauli in enumerate(paulis):
        op1 += WeightedPauliOperator(paulis=[[-coeffs[i], pauli]])
        op1.simplify()
        assert len(op1.paulis) == len(paulis) - (i + 1)
    same = WeightedPauliOperator(paulis=[[0.5, Pauli('IXYZ')], [0.5, Pauli('IXYZ')]])   # :285-298
    assert len(same.paulis) == 1 and len(same.basis) == 1 and same.basis[0][1][0] == 0


@pytest.mark.parametrize('coeffs,counts', [
    ([0.2, 0.6, 0.8, -0.2, -0.6, -0.8], [4, 2, 0]),                                                   # :300-315
    ([0.2 + -0.5j, 0.6 - 0.3j, 0.8 - 0.6j, -0.5 + -0.2j, -0.3 + 0.6j, -0.6 + 0.8j], [6, 2, 0])])     # :317-334
def test_ref_chop(coeffs, counts):
    ori = WeightedPauliOperator.from_list([Pauli(x) for x in SIX], coeffs)
    for threshold, num in zip([0.4, 0.7, 0.9], counts):
        op = ori.copy()
        op1 = op.chop(threshold=threshold, copy=True)
        assert len(op.paulis) == 6 and len(op1.paulis) == num
        op1 = op.chop(threshold=threshold, copy=False)
        assert len(op.paulis) == num and len(op1.paulis) == num


def test_commutation_helpers():
    zz, xx, xi = _term('ZZ', 1.0), _term('XX', 2.0), _term('XI', 0.5)
    assert zz.commute_with(xx) and not zz.commute_with(xi)
    assert zz.anticommute_with(xi) and not zz.anticommute_with(xx)
    assert zz.reorder_paulis() is zz.paulis
