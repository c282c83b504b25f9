"""FermionicOperator helpers (transform, mode elimination / freezing, total particle number,
magnetisation and S^2 -- the aux operators of the eigensolver's chemistry callers) against outputs
of the reference's own class (tests/golden/reference_fermionic.json, made by
tests/golden/make_fermionic_fixtures.py)."""
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, 'golden', 'reference_fermionic.json')))


def arr(rec):
    a = np.zeros(rec['shape'], dtype=complex)
    for *idx, re, im in rec['nonzero']:
        a[tuple(idx)] = complex(re, im)
    return a


@pytest.mark.parametrize('rec', FX['spin_operators'], ids=['%d-modes' % r['modes'] for r in FX['spin_operators']])
def test_spin_observables_equal_reference(rec):
    base = q.FermionicOperator(np.zeros((rec['modes'],) * 2))
    for name in ('total_particle_number', 'total_magnetization', 'total_angular_momentum'):
        f = getattr(base, name)()
        assert np.array_equal(np.asarray(f.h1, dtype=complex), arr(rec[name]['h1'])), name
        assert np.array_equal(np.asarray(f.h2, dtype=complex), arr(rec[name]['h2'])), name
        got = [(p.to_label(), complex(w)) for w, p in f.mapping('jordan_wigner').paulis]
        want = [(lab, complex(*w)) for w, lab in rec[name]['jordan_wigner']]
        assert [g[0] for g in got] == [w[0] for w in want]
        assert np.allclose([g[1] for g in got], [w[1] for w in want], atol=1e-14)


@pytest.mark.parametrize('rec', FX['mode_reduction'], ids=['%d-%s' % (r['modes'], r['list']) for r in FX['mode_reduction']])
def test_mode_elimination_and_freezing_equal_reference(rec):
    h1, h2 = arr(rec['h1']).real, arr(rec['h2']).real
    el = q.FermionicOperator(h1, h2).fermion_mode_elimination(rec['list'])
    assert np.array_equal(el.h1, arr(rec['eliminated']['h1']).real)
    assert np.array_equal(el.h2, arr(rec['eliminated']['h2']).real)
    fr, shift = q.FermionicOperator(h1, h2).fermion_mode_freezing(rec['list'])
    assert np.allclose(fr.h1, arr(rec['frozen']['h1']).real, rtol=1e-13, atol=1e-13)
    assert np.array_equal(fr.h2, arr(rec['frozen']['h2']).real)
    assert shift == pytest.approx(rec['frozen']['energy_shift'], rel=1e-13, abs=1e-13)


@pytest.mark.parametrize('rec', FX['transform'], ids=['%d-modes' % r['modes'] for r in FX['transform']])
def test_transform_equals_reference(rec):
    f = q.FermionicOperator(arr(rec['h1']), arr(rec['h2']))
    f.transform(arr(rec['unitary']))
    assert np.allclose(f.h1, arr(rec['h1_out']), atol=1e-12)
    assert np.allclose(f.h2, arr(rec['h2_out']), atol=1e-12)
    g = q.FermionicOperator(arr(rec['h1']), arr(rec['h2']))
    assert g != f and g == q.FermionicOperator(arr(rec['h1']), arr(rec['h2']))


def test_freezing_preserves_the_sector_energy():
    """Independent of the fixtures: with mode 0 frozen (occupied), <H> over states that have mode 0
    occupied equals <H_frozen> + shift."""
    from oracle import pauli_oracle as po
    rng = np.random.default_rng(9)
    n = 4
    h1 = rng.standard_normal((n, n))
    h1 = h1 + h1.T
    h2 = np.zeros((n,) * 4)
    for i in range(n):
        for l in range(n):
            h2[i, i, l, l] = rng.standard_normal()
    full = q.FermionicOperator(h1, h2).mapping('jordan_wigner')
    frozen, shift = q.FermionicOperator(h1, h2).fermion_mode_freezing([0])
    red = frozen.mapping('jordan_wigner')
    mf = po.ref_to_spmatrix([(complex(w), p.to_label()) for w, p in full.paulis]).toarray()
    mr = po.ref_to_spmatrix([(complex(w), p.to_label()) for w, p in red.paulis]).toarray()
    occupied = [i for i in range(1 << n) if i & 1]          # qubit 0 = mode 0 occupied
    sub = mf[np.ix_(occupied, occupied)]
    assert np.allclose(np.linalg.eigvalsh(sub), np.linalg.eigvalsh(mr) + shift, atol=1e-10)
