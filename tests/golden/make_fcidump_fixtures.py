"""Runs the REFERENCE'S OWN FCIDump parser / dumper / driver and QMolecule spin-orbital conversion
(qiskit/chemistry/drivers/fcidumpd/{parser,dumper,fcidumpdriver}.py, qiskit/chemistry/qmolecule.py)
on the reference's three FCIDump test files (H2 and LiH restricted, OH unrestricted) and records
inputs and outputs in tests/golden/reference_fcidump.json.

    python tests/golden/make_fcidump_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py (terra stubbed, package __init__ files bypassed;
h5py, which QMolecule only needs for save/load, is a dummy module).
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402

TESTDIR = '/root/reference/test/chemistry'


def sparse(a):
    if a is None:
        return None
    a = np.asarray(a)
    return {'shape': list(a.shape), 'nonzero': [[int(v) for v in idx] + [float(a[idx])] for idx in zip(*np.nonzero(a))]}


def digest(a):
    """Compact fingerprint of a large array: count of non-zeros and three weighted sums."""
    a = np.asarray(a)
    idx = np.indices(a.shape)
    w = sum((k + 1) * idx[k] for k in range(a.ndim)) + 1.0
    return {'shape': list(a.shape), 'count': int(np.count_nonzero(a)), 'sum': float(a.sum()),
            'abs_sum': float(np.abs(a).sum()), 'weighted': float((a * w).sum())}


def main():
    R = mrf.load_reference()
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))

    def synth(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
        parent, _, leaf = name.rpartition('.')
        setattr(sys.modules[parent], leaf, m)
        return m

    synth('qiskit.chemistry.drivers', mrf.REF + '/chemistry/drivers')
    synth('qiskit.chemistry.drivers.fcidumpd', mrf.REF + '/chemistry/drivers/fcidumpd')
    parser = importlib.import_module('qiskit.chemistry.drivers.fcidumpd.parser')
    dumper = importlib.import_module('qiskit.chemistry.drivers.fcidumpd.dumper')
    QMolecule = importlib.import_module('qiskit.chemistry.qmolecule').QMolecule
    out = {'note': 'outputs of the reference FCIDump parser/dumper and QMolecule.onee/twoe_to_spin; see make_fcidump_fixtures.py',
           'cases': []}
    for name in ('h2', 'lih', 'oh'):
        path = os.path.join(TESTDIR, 'test_driver_fcidump_%s.fcidump' % name)
        data = parser.parse(path)
        rec = {'name': name, 'text': open(path).read(),
               'header': {k: data[k] for k in ('NORB', 'NELEC', 'MS2', 'ISYM', 'ORBSYM', 'IPRTIM', 'INT', 'MEMORY',
                                               'CORE', 'MAXIT', 'THR', 'THRRES', 'NROOT')},
               'ecore': data.get('ecore')}
        for key in ('hij', 'hij_b', 'hijkl', 'hijkl_ba', 'hijkl_bb'):
            rec[key] = sparse(data[key])
        # what FCIDumpDriver.run() derives (fcidumpdriver.py:66-88)
        mult = data.get('MS2', 0) + 1
        nbeta = (data['NELEC'] - (mult - 1)) // 2
        rec['multiplicity'], rec['num_beta'], rec['num_alpha'] = mult, nbeta, data['NELEC'] - nbeta
        h1 = QMolecule.onee_to_spin(data['hij'], data['hij_b'])
        h2 = QMolecule.twoe_to_spin(data['hijkl'], data['hijkl_bb'], data['hijkl_ba'])
        rec['one_body_integrals'] = sparse(h1)
        rec['two_body_integrals'] = digest(h2)
        with tempfile.TemporaryDirectory() as d:
            outp = os.path.join(d, 'out.fcidump')
            dumper.dump(outp, data['NORB'], data['NELEC'], (data['hij'], data['hij_b']),
                        (data['hijkl'], data['hijkl_ba'], data['hijkl_bb']), data['ecore'], ms2=data['MS2'])
            rec['dumped'] = open(outp).read()
            back = parser.parse(outp)
            rec['roundtrip_ok'] = bool(np.allclose(back['hijkl'], data['hijkl']) and np.allclose(back['hij'], data['hij']))
        if name in ('h2', 'lih'):
            op = R['FermionicOperator'](h1, h2).mapping('jordan_wigner')
            rec['jordan_wigner_terms'] = len(op.paulis)
            if name == 'h2':
                rec['jordan_wigner'] = [[[complex(w).real, complex(w).imag], p.to_label()] for w, p in op.paulis]
        out['cases'].append(rec)
    path = os.path.join(HERE, 'reference_fcidump.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')
    for c in out['cases']:
        print(c['name'], c['header']['NORB'], c['header']['NELEC'], c['header']['MS2'], 'uhf' if c['hij_b'] else 'rhf',
              len(c['hijkl']['nonzero']), c['roundtrip_ok'], c.get('jordan_wigner_terms'))


if __name__ == '__main__':
    main()
