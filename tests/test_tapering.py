"""Z2Symmetries (qiskit_aqua_b200/tapering.py) against outputs of the reference's own class
(tests/golden/reference_tapering.json, made by tests/golden/make_tapering_fixtures.py): the
FCIDump -> parity mapping -> two-qubit reduction chain that produces the reference's golden 2-qubit
H2 operator, symmetry discovery, and tapering in every sector."""
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q
from qiskit_aqua_b200.tapering import Z2Symmetries, kernel_F2
from oracle import pauli_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, 'golden', 'reference_tapering.json')))
CHEM = json.load(open(os.path.join(HERE, 'golden', 'chemistry_integrals.json')))
G1 = {'II': -1.052373245772859, 'IZ': 0.39793742484318045, 'ZI': -0.39793742484318045,
      'ZZ': -0.01128010425623538, 'XX': 0.18093119978423156}     # test/aqua/test_numpy_eigen_solver.py:26-37


def same_paulis(op, want, tol=1e-12):
    got = [(p.to_label(), complex(w)) for w, p in op.paulis]
    want = [(lab, complex(*w)) for w, lab in want]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=tol, atol=tol)


def parity_op(key):
    mol = q.QMolecule()
    h1 = q.QMolecule.onee_to_spin(np.array(CHEM[key]['mo_onee']))
    h2 = q.QMolecule.twoe_to_spin(np.array(CHEM[key]['mo_eri']))
    del mol
    return q.FermionicOperator(h1, h2).mapping('parity')


def test_h2_two_qubit_reduction_is_the_golden_operator():
    rec = FX['two_qubit_reduction'][0]
    red = Z2Symmetries.two_qubit_reduction(parity_op('G10_h2_fcidump'), rec['num_particles'])
    assert red.num_qubits == 2
    same_paulis(red, rec['reduced'])
    assert red.z2_symmetries.tapering_values == rec['tapering_values']
    assert red.z2_symmetries.sq_list == rec['sq_list']
    got = {p.to_label(): complex(w).real for w, p in red.paulis}
    assert set(got) == set(G1)
    for lab, val in G1.items():                   # the operator every reference eigensolver test uses
        assert got[lab] == pytest.approx(val, abs=1e-9)


def test_lih_two_qubit_reduction_matches_reference():
    rec = FX['two_qubit_reduction'][1]
    op = parity_op('G11_lih_npz')
    assert len(op.paulis) == rec['parity_terms']
    red = Z2Symmetries.two_qubit_reduction(op, rec['num_particles'])
    assert red.num_qubits == rec['num_qubits'] == 10
    assert len(red.paulis) == rec['reduced_terms']
    d = {p.to_label(): complex(w) for w, p in red.paulis}
    for w, lab in rec['reduced_sample']:
        assert d[lab] == pytest.approx(complex(*w), rel=1e-11, abs=1e-13)
    assert sum(d.values()) == pytest.approx(complex(*rec['reduced_weight_sum']), rel=1e-11)
    assert sum(abs(v) for v in d.values()) == pytest.approx(rec['reduced_abs_sum'], rel=1e-11)


@pytest.mark.parametrize('rec', FX['find_and_taper'], ids=[r['name'] for r in FX['find_and_taper']])
def test_find_symmetries_and_taper_match_reference(rec):
    op = q.WeightedPauliOperator(paulis=[[complex(*w), q.Pauli(l)] for w, l in rec['terms']])
    z2 = Z2Symmetries.find_Z2_symmetries(op)
    assert [s.to_label() for s in z2.symmetries] == rec['symmetries']
    assert [s.to_label() for s in z2.sq_paulis] == rec['sq_paulis']
    assert z2.sq_list == rec['sq_list']
    if rec['sectors'] is None:
        assert z2.is_empty()
        with pytest.raises(q.AquaError):
            z2.taper(op)
        return
    sectors = z2.taper(op)
    assert len(sectors) == len(rec['sectors'])
    for got, want in zip(sectors, rec['sectors']):
        assert got.z2_symmetries.tapering_values == want['tapering_values']
        same_paulis(got, want['paulis'])
    # independent of the fixtures: the sectors together carry the whole spectrum
    if len(set(z2.sq_list)) == len(z2.sq_list):
        full = np.sort(np.linalg.eigvalsh(po.ref_to_spmatrix([(complex(*w), l) for w, l in rec['terms']]).toarray()))
        parts = []
        for t in sectors:
            if t.is_empty():
                parts.append(np.zeros(1 << (op.num_qubits - len(z2.sq_list))))
            else:
                m = po.ref_to_spmatrix([(complex(w), p.to_label()) for w, p in t.paulis]).toarray()
                parts.append(np.linalg.eigvalsh(m))
        assert np.allclose(np.sort(np.concatenate(parts)), full, atol=1e-9)


def test_kernel_and_errors():
    # commuting with X0 X1 and Z0 Z1: the kernel of {XX, ZZ} over 2 qubits is spanned by XX, ZZ
    ker = kernel_F2([0b0011, 0b1100], 4)
    assert sorted(ker) == [0b0011, 0b1100]
    with pytest.raises(q.AquaError):
        Z2Symmetries([q.Pauli('ZZ')], [], [])
    with pytest.raises(q.AquaError):
        Z2Symmetries([q.Pauli('ZZ')], [q.Pauli('IX')], [0], tapering_values=[1, 1])
    z2 = Z2Symmetries([q.Pauli('ZZ')], [q.Pauli('IX')], [0], [1])
    with pytest.raises(q.AquaError):              # XI does not commute with the symmetry ZZ
        z2.consistent_tapering(q.WeightedPauliOperator(paulis=[[1.0, q.Pauli('XI')]]))
    aux = z2.consistent_tapering(q.WeightedPauliOperator(paulis=[[2.0, q.Pauli('XX')], [1.0, q.Pauli('ZZ')]]))
    assert aux.num_qubits == 1
    assert 'Tapering values' in str(z2)


@pytest.mark.gpu
def test_tapered_h2_spectrum_on_device():
    """G1 through the whole chain: integrals -> parity -> two-qubit reduction -> device eigensolver."""
    red = Z2Symmetries.two_qubit_reduction(parity_op('G10_h2_fcidump'), 2)
    res = q.NumPyEigensolver(red, k=4).run()
    assert np.allclose(np.real(res.eigenvalues), [-1.85727503, -1.24458455, -0.88272215, -0.22491125], atol=1e-7)


@pytest.mark.gpu
def test_tapered_lih_ground_state_on_device():
    """10-qubit tapered LiH has the same ground energy as the 12-qubit Hamiltonian (G11)."""
    red = Z2Symmetries.two_qubit_reduction(parity_op('G11_lih_npz'), [2, 2])
    res = q.NumPyMinimumEigensolver(red).run()
    assert res.eigenvalue.real == pytest.approx(-8.87487309, abs=1e-6)
