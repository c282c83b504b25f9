"""Runs the REFERENCE'S OWN FermionicOperator helpers (qiskit/chemistry/fermionic_operator.py:
transform, fermion_mode_elimination, fermion_mode_freezing, total_particle_number,
total_magnetization, total_angular_momentum) on seeded integrals and records the results in
tests/golden/reference_fermionic.json.

    python tests/golden/make_fermionic_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py (terra stubbed, package __init__ files bypassed).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402


def carr(a):
    a = np.asarray(a)
    nz = list(zip(*np.nonzero(a)))
    return {'shape': list(a.shape), 'nonzero': [[int(v) for v in idx] + [float(np.real(a[idx])), float(np.imag(a[idx]))] for idx in nz]}


def plist(op):
    return [[[complex(w).real, complex(w).imag], p.to_label()] for w, p in op.paulis]


def main():
    R = mrf.load_reference()
    FO = R['FermionicOperator']
    out = {'note': 'outputs of the reference FermionicOperator helpers (terra Pauli stubbed); see make_fermionic_fixtures.py',
           'spin_operators': [], 'mode_reduction': [], 'transform': []}
    for modes in (2, 4, 6):
        base = FO(np.zeros((modes, modes)))
        rec = {'modes': modes}
        for name in ('total_particle_number', 'total_magnetization', 'total_angular_momentum'):
            f = getattr(base, name)()
            rec[name] = {'h1': carr(f.h1), 'h2': carr(f.h2), 'jordan_wigner': plist(f.mapping('jordan_wigner'))}
        out['spin_operators'].append(rec)
    r = np.random.default_rng(606)
    for modes, nnz, lists in ((6, 60, ([0, 3], [1], [2, 5], [4, 0, 1])), (4, 256, ([0, 2], [3]))):
        h1 = r.standard_normal((modes, modes))
        h1 = h1 + h1.T
        h2 = np.zeros((modes,) * 4)
        if nnz >= modes ** 4:
            h2 = r.standard_normal((modes,) * 4)
        else:
            for _ in range(nnz):
                i, j, k, m = r.integers(0, modes, size=4)
                h2[i, j, k, m] = r.standard_normal()
            for i in range(modes):            # make sure the pair terms that freezing folds are present
                for l in range(modes):
                    h2[i, i, l, l] = r.standard_normal()
                    h2[i, l, l, i] = r.standard_normal()
        for modes_out in lists:
            el = FO(h1, h2).fermion_mode_elimination(modes_out)
            fr, shift = FO(h1, h2).fermion_mode_freezing(modes_out)
            out['mode_reduction'].append({'modes': modes, 'h1': carr(h1), 'h2': carr(h2), 'list': list(modes_out),
                                          'eliminated': {'h1': carr(el.h1), 'h2': carr(el.h2)},
                                          'frozen': {'h1': carr(fr.h1), 'h2': carr(fr.h2), 'energy_shift': float(shift)}})
    for modes in (3, 4):
        h1 = r.standard_normal((modes, modes)) + 1j * r.standard_normal((modes, modes))
        h2 = r.standard_normal((modes,) * 4)
        q, _ = np.linalg.qr(r.standard_normal((modes, modes)) + 1j * r.standard_normal((modes, modes)))
        f = FO(h1, h2)
        f.transform(q)
        out['transform'].append({'modes': modes, 'h1': carr(h1), 'h2': carr(h2), 'unitary': carr(q),
                                 'h1_out': carr(f.h1), 'h2_out': carr(f.h2)})
    path = os.path.join(HERE, 'reference_fermionic.json')
    with open(path, 'w') as fh:
        json.dump(out, fh, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
