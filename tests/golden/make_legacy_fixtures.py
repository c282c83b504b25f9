"""Runs the REFERENCE'S OWN MatrixOperator / op_converter (qiskit/aqua/operators/legacy/
{matrix_operator,op_converter}.py) on seeded matrices and Pauli sums and records the results in
tests/golden/reference_legacy.json.

    python tests/golden/make_legacy_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py (terra stubbed, package __init__ files bypassed).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402
import terra_stub  # noqa: E402


def cplx(v):
    v = complex(v)
    return [v.real, v.imag]


def cmat(m):
    m = np.asarray(m.toarray() if hasattr(m, 'toarray') else m, dtype=complex)
    return {'re': m.real.tolist(), 'im': m.imag.tolist()}


def main():
    R = mrf.load_reference()
    WPO, MO, conv, Pauli = R['WPO'], R['MatrixOperator'], R['op_converter'], terra_stub.Pauli
    out = {'note': 'outputs of the reference MatrixOperator / op_converter (terra Pauli stubbed); see make_legacy_fixtures.py',
           'decompose': [], 'pauli_sums': []}
    r = np.random.default_rng(2024)
    mats = []
    for n in (1, 2, 3):
        mats.append(('generic%d' % n, r.standard_normal((1 << n, 1 << n)) + 1j * r.standard_normal((1 << n, 1 << n))))
    a = r.standard_normal((8, 8)) + 1j * r.standard_normal((8, 8))
    mats.append(('hermitian3', a + a.conj().T))
    mats.append(('diagonal3', np.diag(r.standard_normal(8))))
    mats.append(('complexdiag2', np.diag(r.standard_normal(4) + 1j * r.standard_normal(4))))
    sparse = np.zeros((8, 8), dtype=complex)
    for _ in range(9):
        i, j = r.integers(0, 8, size=2)
        sparse[i, j] = r.standard_normal()
    mats.append(('sparse3', sparse))
    mats.append(('real_symmetric4', (lambda b: b + b.T)(r.standard_normal((16, 16)))))
    for name, m in mats:
        mo = MO(matrix=m)
        op = conv.to_weighted_pauli_operator(mo)
        psi = r.standard_normal(m.shape[0]) + 1j * r.standard_normal(m.shape[0])
        psi /= np.linalg.norm(psi)
        chopped = mo.chop(0.3, copy=True)
        out['decompose'].append({
            'name': name, 'matrix': cmat(m), 'num_qubits': int(mo.num_qubits), 'str': str(mo),
            'is_diagonal': mo.dia_matrix is not None,
            'matrix_property': cmat(mo.matrix) if mo.dia_matrix is None else cmat(np.asarray(mo.matrix).reshape(1, -1)),
            'paulis': [[cplx(w), p.to_label()] for w, p in op.paulis],
            'psi': cmat(psi.reshape(1, -1)), 'expectation': cplx(mo.evaluate_with_statevector(psi)[0]),
            'chop_0.3': cmat(chopped.dense_matrix), 'chop_0.3_empty': bool(chopped.is_empty())})
    # Pauli sums -> MatrixOperator and the MatrixOperator algebra
    sums = [('h2', [(-1.052373245772859, 'II'), (0.39793742484318045, 'ZI'), (-0.39793742484318045, 'IZ'),
                    (-0.01128010425623538, 'ZZ'), (0.18093119978423156, 'XX')]),
            ('complex3', [(0.5 + 0.25j, 'XYZ'), (-1.0, 'ZZI'), (0.3j, 'IYI'), (0.8, 'III')]),
            ('diag3', [(1.0, 'ZIZ'), (-0.5, 'IZI'), (0.25, 'III')])]
    other = [(0.7, 'XI'), (-0.2, 'ZZ'), (0.1j, 'YX')]
    for name, terms in sums:
        op = WPO(paulis=[[w, Pauli(l)] for w, l in terms])
        mo = conv.to_matrix_operator(op)
        rec = {'name': name, 'terms': [[cplx(w), l] for w, l in terms], 'dense': cmat(mo.dense_matrix),
               'is_diagonal': mo.dia_matrix is not None, 'num_qubits': int(mo.num_qubits),
               'back': [[cplx(w), p.to_label()] for w, p in conv.to_weighted_pauli_operator(mo).paulis]}
        if name == 'h2':
            mo2 = conv.to_matrix_operator(WPO(paulis=[[w, Pauli(l)] for w, l in other]))
            rec['other'] = [[cplx(w), l] for w, l in other]
            rec['add'], rec['sub'], rec['mul'], rec['neg'] = (cmat((mo + mo2).dense_matrix), cmat((mo - mo2).dense_matrix),
                                                              cmat((mo * mo2).dense_matrix), cmat((-mo).dense_matrix))
            # (the reference's __eq__ -- np.all(csr == csr) -- raises with current numpy/scipy; not recorded)
        out['pauli_sums'].append(rec)
    path = os.path.join(HERE, 'reference_legacy.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')
    for d in out['decompose']:
        print(d['name'], d['num_qubits'], d['is_diagonal'], len(d['paulis']), [p[1] for p in d['paulis'][:6]])


if __name__ == '__main__':
    main()
