"""Runs the REFERENCE'S OWN FermionicOperator.mapping (qiskit/chemistry/fermionic_operator.py) for
the jordan_wigner, parity and bravyi_kitaev encodings on seeded integrals and records the resulting
qubit operators in tests/golden/reference_mappings.json.

    python tests/golden/make_mapping_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py (terra stubbed, package __init__ files bypassed).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402


def cplx(v):
    v = complex(v)
    return [v.real, v.imag]


def main():
    R = mrf.load_reference()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import pauli_oracle as po
    chem = json.load(open(os.path.join(HERE, 'chemistry_integrals.json')))['G10_h2_fcidump']
    cases = [('h2_fcidump', po.onee_to_spin(np.array(chem['mo_onee'])), po.twoe_to_spin(np.array(chem['mo_eri'])))]
    for modes, seed, nnz in ((3, 1, 6), (5, 7, 10), (6, 2, 14), (8, 3, 12)):
        r = np.random.default_rng(seed)
        g1 = r.standard_normal((modes, modes))
        g1 = g1 + g1.T
        g2 = np.zeros((modes,) * 4)
        for _ in range(nnz):
            i, j, k, m = r.integers(0, modes, size=4)
            g2[i, j, k, m] = r.standard_normal()
        cases.append(('random%d' % modes, g1, g2))
    out = {'note': 'outputs of the reference FermionicOperator.mapping (terra Pauli stubbed); see make_mapping_fixtures.py',
           'cases': []}
    for name, h1, h2 in cases:
        rec = {'name': name, 'modes': int(h1.shape[0]), 'h1': h1.tolist(),
               'h2_nonzero': [[int(v) for v in idx] + [float(h2[idx])] for idx in zip(*np.nonzero(h2))], 'mappings': {}}
        for map_type in ('jordan_wigner', 'parity', 'bravyi_kitaev'):
            op = R['FermionicOperator'](h1, h2).mapping(map_type)
            rec['mappings'][map_type] = [[cplx(w), p.to_label()] for w, p in op.paulis]
        out['cases'].append(rec)
    path = os.path.join(HERE, 'reference_mappings.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
