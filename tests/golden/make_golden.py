"""Regenerates tests/golden/*.json from the reference's own test fixtures.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The outputs are small data fixtures (molecular integrals and the reference's known answers);
no reference source code is copied.  Sources:
  * test/chemistry/test_driver_fcidump_h2.fcidump  (G10: NORB=2 FCIDump text, 8 integrals)
  * test/chemistry/test_driver_fcidump_lih.npz     (G11: mo_onee 6x6, mo_eri 6^4)
  * known answers quoted in SURVEY.md section 8c with their reference test file:line.
"""
import itertools
import json
import os

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_fcidump(path):
    """FCIDump text -> (mo_onee, mo_eri in chemists' (ij|kl) with the 8-fold symmetry filled,
    nuclear repulsion).  Follows qiskit/chemistry/drivers/fcidumpd/parser.py:30-199."""
    txt = open(path).read()
    head, body = txt.split('/', 1) if '&END' not in txt else txt.split('&END', 1)
    norb = int(head.upper().split('NORB=')[1].split(',')[0])
    h1 = np.zeros((norb, norb))
    h2 = np.zeros((norb,) * 4)
    enuc = 0.0
    for line in body.strip().splitlines():
        parts = line.split()
        if len(parts) != 5:
            continue
        v = float(parts[0])
        i, a, j, b = (int(t) for t in parts[1:])
        if i == a == j == b == 0:
            enuc = v
        elif j == b == 0:
            h1[i - 1, a - 1] = v
            h1[a - 1, i - 1] = v
        else:
            i, a, j, b = i - 1, a - 1, j - 1, b - 1
            for (p, q), (r, s) in itertools.product(((i, a), (a, i)), ((j, b), (b, j))):
                h2[p, q, r, s] = v
                h2[r, s, p, q] = v
    return h1, h2, enuc


def main():
    h1, h2, enuc = parse_fcidump(os.path.join(REF, 'test/chemistry/test_driver_fcidump_h2.fcidump'))
    lih = np.load(os.path.join(REF, 'test/chemistry/test_driver_fcidump_lih.npz'))
    out = {
        'G10_h2_fcidump': {
            'source': 'test/chemistry/test_driver_fcidump_h2.fcidump',
            'mo_onee': h1.tolist(), 'mo_eri': h2.tolist(), 'nuclear_repulsion': enuc,
            'expect_num_qubits': 4, 'expect_num_paulis': 15,          # test_core_hamiltonian.py:74-86
            'expect_min_eigenvalue': -1.857275027031588,              # test_end2end_with_vqe.py:47
        },
        'G11_lih_npz': {
            'source': 'test/chemistry/test_driver_fcidump_lih.npz',
            'mo_onee': np.asarray(lih['mo_onee']).tolist(),
            'mo_eri': np.asarray(lih['mo_eri']).tolist(),
            'nuclear_repulsion': 0.9924,                                # test_driver_fcidump.py:152
            'expect_total_energy_3dp': -7.882,                          # test_driver_methods_gsc.py:32
        },
    }
    with open(os.path.join(HERE, 'chemistry_integrals.json'), 'w') as f:
        json.dump(out, f)
    print('wrote chemistry_integrals.json', os.path.getsize(os.path.join(HERE, 'chemistry_integrals.json')))


if __name__ == '__main__':
    main()
