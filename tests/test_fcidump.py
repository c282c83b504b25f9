"""FCIDump reader / writer / driver and spin-orbital conversion (qiskit_aqua_b200/fcidump.py)
against outputs of the reference's own parser, dumper and QMolecule on the reference's three
FCIDump test files (tests/golden/reference_fcidump.json, made by make_fcidump_fixtures.py), and the
file -> Hamiltonian -> ground-state chain on the device."""
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q
from qiskit_aqua_b200 import fcidump

HERE = os.path.dirname(os.path.abspath(__file__))
FX = {c['name']: c for c in json.load(open(os.path.join(HERE, 'golden', 'reference_fcidump.json')))['cases']}


def dense(rec):
    if rec is None:
        return None
    a = np.zeros(rec['shape'])
    for *idx, v in rec['nonzero']:
        a[tuple(idx)] = v
    return a


def digest(a):
    idx = np.indices(a.shape)
    w = sum((k + 1) * idx[k] for k in range(a.ndim)) + 1.0
    return {'shape': list(a.shape), 'count': int(np.count_nonzero(a)), 'sum': float(a.sum()),
            'abs_sum': float(np.abs(a).sum()), 'weighted': float((a * w).sum())}


def write(tmp_path, name):
    path = str(tmp_path / (name + '.fcidump'))
    with open(path, 'w') as f:
        f.write(FX[name]['text'])
    return path


@pytest.mark.parametrize('name', ['h2', 'lih', 'oh'])
def test_parse_equals_reference(tmp_path, name):
    rec = FX[name]
    data = fcidump.parse(write(tmp_path, name))
    for key, want in rec['header'].items():
        assert data[key] == want, key
    assert data.get('ecore') == rec['ecore']
    for key in ('hij', 'hij_b', 'hijkl', 'hijkl_ba', 'hijkl_bb'):
        want = dense(rec[key])
        if want is None:
            assert data[key] is None
        else:
            assert np.array_equal(data[key], want), key


@pytest.mark.parametrize('name', ['h2', 'lih', 'oh'])
def test_driver_and_spin_integrals_equal_reference(tmp_path, name):
    rec = FX[name]
    mol = fcidump.FCIDumpDriver(write(tmp_path, name)).run()
    assert (mol.num_orbitals, mol.multiplicity, mol.num_alpha, mol.num_beta) == \
        (rec['header']['NORB'], rec['multiplicity'], rec['num_alpha'], rec['num_beta'])
    assert mol.nuclear_repulsion_energy == rec['ecore']
    assert np.array_equal(mol.one_body_integrals, dense(rec['one_body_integrals']))
    got, want = digest(mol.two_body_integrals), rec['two_body_integrals']
    assert got['shape'] == want['shape'] and got['count'] == want['count']
    for key in ('sum', 'abs_sum', 'weighted'):
        assert got[key] == pytest.approx(want[key], rel=1e-12, abs=1e-12)


@pytest.mark.parametrize('name', ['h2', 'lih', 'oh'])
def test_dump_equals_reference_and_round_trips(tmp_path, name):
    rec = FX[name]
    mol = fcidump.FCIDumpDriver(write(tmp_path, name)).run()
    out = str(tmp_path / 'out.fcidump')
    fcidump.FCIDumpDriver.dump(mol, out)
    assert open(out).read() == rec['dumped']
    back = fcidump.parse(out)
    assert np.allclose(back['hij'], mol.mo_onee_ints, atol=1e-15)
    assert np.allclose(back['hijkl'], mol.mo_eri_ints, atol=1e-15)
    if mol.mo_eri_ints_ba is not None:
        assert np.allclose(back['hijkl_ba'], mol.mo_eri_ints_ba, atol=1e-15)
        assert np.allclose(back['hijkl_bb'], mol.mo_eri_ints_bb, atol=1e-15)
        assert np.allclose(back['hij_b'], mol.mo_onee_ints_b, atol=1e-15)


def test_h2_chain_equals_reference_operator(tmp_path):
    """FCIDump -> spin integrals -> Jordan-Wigner: the reference's 15-term H2 operator
    (test/chemistry/test_core_hamiltonian.py:54-58, 74-86), term by term."""
    mol = fcidump.FCIDumpDriver(write(tmp_path, 'h2')).run()
    op = q.FermionicOperator(h1=mol.one_body_integrals, h2=mol.two_body_integrals).mapping('jordan_wigner')
    got = [(p.to_label(), complex(w)) for w, p in op.paulis]
    want = [(lab, complex(*w)) for w, lab in FX['h2']['jordan_wigner']]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=1e-13, atol=1e-15)
    assert mol.nuclear_repulsion_energy == pytest.approx(0.71997, abs=1e-5)


def test_parse_errors(tmp_path):
    with pytest.raises(fcidump.QiskitChemistryError):
        fcidump.parse(str(tmp_path / 'missing.fcidump'))
    p = str(tmp_path / 'bad.fcidump')
    open(p, 'w').write('&FCI NELEC= 2,MS2=0,\n /\n 1.0 1 1 0 0\n')
    with pytest.raises(fcidump.QiskitChemistryError):          # NORB missing
        fcidump.parse(p)
    open(p, 'w').write('&FCI NORB= 2,NELEC= 2,MS2=0,\n /\n 1.0 1 3 0 0\n')
    with pytest.raises(fcidump.QiskitChemistryError):          # index out of range
        fcidump.parse(p)
    with pytest.raises(fcidump.QiskitChemistryError):
        fcidump.FCIDumpDriver(123)
    assert issubclass(fcidump.QiskitChemistryError, q.AquaError)


@pytest.mark.gpu
def test_h2_file_to_ground_energy_on_device(tmp_path):
    """test/chemistry/test_end2end_with_vqe.py:47: -1.857275027031588 electronic, + 0.71997 nuclear."""
    mol = fcidump.FCIDumpDriver(write(tmp_path, 'h2')).run()
    op = q.FermionicOperator(h1=mol.one_body_integrals, h2=mol.two_body_integrals).mapping('jordan_wigner')
    res = q.NumPyMinimumEigensolver(op).run()
    assert res.eigenvalue.real == pytest.approx(-1.857275030202, abs=1e-8)
    assert res.eigenvalue.real + mol.nuclear_repulsion_energy == pytest.approx(-1.137306, abs=1e-5)


@pytest.mark.gpu
def test_lih_file_to_ground_energy_on_device(tmp_path):
    """G11: the 12-qubit LiH Hamiltonian from the FCIDump text; total energy -7.882
    (test/chemistry/test_driver_methods_gsc.py:32), through the device Lanczos branch."""
    mol = fcidump.FCIDumpDriver(write(tmp_path, 'lih')).run()
    op = q.FermionicOperator(h1=mol.one_body_integrals, h2=mol.two_body_integrals).mapping('jordan_wigner')
    assert op.num_qubits == 12 and len(op.paulis) == FX['lih']['jordan_wigner_terms']
    res = q.NumPyMinimumEigensolver(op).run()
    assert res.eigenvalue.real == pytest.approx(-8.87487309, abs=1e-6)
    assert res.eigenvalue.real + mol.nuclear_repulsion_energy == pytest.approx(-7.8825, abs=1e-3)
