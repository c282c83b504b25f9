"""FermionicOperator.mapping (jordan_wigner / parity / bravyi_kitaev) against outputs of the
reference's own qiskit/chemistry/fermionic_operator.py (tests/golden/reference_mappings.json,
produced by tests/golden/make_mapping_fixtures.py)."""
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q
from oracle import pauli_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, 'golden', 'reference_mappings.json')))
MAPS = ('jordan_wigner', 'parity', 'bravyi_kitaev')


def integrals(case):
    n = case['modes']
    h2 = np.zeros((n,) * 4)
    for i, j, k, m, v in case['h2_nonzero']:
        h2[i, j, k, m] = v
    return np.array(case['h1']), h2


@pytest.mark.parametrize('case', FX['cases'], ids=[c['name'] for c in FX['cases']])
@pytest.mark.parametrize('map_type', MAPS)
def test_mapping_equals_reference_run(case, map_type):
    h1, h2 = integrals(case)
    op = q.FermionicOperator(h1, h2).mapping(map_type)
    got = [(p.to_label(), complex(w)) for w, p in op.paulis]
    want = [(lab, complex(*w)) for w, lab in case['mappings'][map_type]]
    assert [g[0] for g in got] == [w[0] for w in want]              # same terms, same order
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize('case', FX['cases'][:4], ids=[c['name'] for c in FX['cases'][:4]])
def test_encodings_are_isospectral(case):
    """Independent of the fixtures: the three encodings are unitarily equivalent."""
    h1, h2 = integrals(case)
    invariants = []
    for map_type in MAPS:
        op = q.FermionicOperator(h1, h2).mapping(map_type)
        m = po.ref_to_spmatrix([(complex(w), p.to_label()) for w, p in op.paulis]).toarray()
        # similarity invariants tr(M^k) (the random h2 is not Hermitian, so sorted complex
        # eigenvalues of degenerate pairs are not a stable comparison)
        mk, tr = np.eye(len(m)), []
        for _ in range(4):
            mk = mk @ m
            tr.append(np.trace(mk))
        invariants.append(np.array(tr))
    scale = np.abs(invariants[0]) + 1.0
    assert np.all(np.abs(invariants[0] - invariants[1]) <= 1e-10 * scale)
    assert np.all(np.abs(invariants[0] - invariants[2]) <= 1e-10 * scale)


def test_unknown_mapping_raises():
    with pytest.raises(ValueError):
        q.FermionicOperator(np.eye(2)).mapping('ternary_tree')


@pytest.mark.gpu
@pytest.mark.parametrize('map_type', MAPS)
def test_h2_ground_energy_on_device(map_type):
    """test/chemistry/test_end2end_with_vqe.py:47 / G10: -1.857275... for every encoding."""
    h1, h2 = integrals(FX['cases'][0])
    op = q.FermionicOperator(h1, h2).mapping(map_type)
    assert op.num_qubits == 4 and len(op.paulis) == 15
    res = q.NumPyMinimumEigensolver(op).run()
    assert res.eigenvalue.real == pytest.approx(-1.857275030202, abs=1e-8)
