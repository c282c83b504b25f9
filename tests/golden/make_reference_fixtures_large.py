"""Reference-run fixtures at sizes the TILED kernel serves (11-13 qubits; reference_run.json stops at 6, where
the streaming kernel runs).  Executes the reference's own WeightedPauliOperator / op_converter / MatrixOperator
CSR path and its NumPyEigensolver (scipy eigs, also with complex weights = ARPACK's non-symmetric Arnoldi)
exactly like make_reference_fixtures.py (same loader, terra's Pauli stubbed) and records

    evaluate_with_statevector, sampled rows of H.psi with its norm, the k lowest eigenvalues

into tests/golden/reference_run_large.json.  States are regenerated from their seed by the tests
(numpy Generator streams are stable); a checksum of psi guards that.

    python tests/golden/make_reference_fixtures_large.py        (needs /root/reference; CPU only)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import terra_stub  # noqa: E402
from make_reference_fixtures import cplx, load_reference  # noqa: E402

CASES = [  # (qubits, terms, complex weights, seed, eigenpairs)
    (11, 60, False, 1101, 2),
    (12, 90, True, 1202, 2),
    (13, 120, False, 1303, 1),
    (12, 40, True, 1204, 1),
    (11, 200, False, 1105, 3),
]


def make_state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return psi / np.linalg.norm(psi)


def make_terms(n, T, cplx_w, seed):
    rng = np.random.default_rng(seed + 7)
    labels = sorted({''.join(rng.choice(list('IXYZ'), p=[0.55, 0.15, 0.15, 0.15]) for _ in range(n)) for _ in range(T)})
    w = rng.standard_normal(len(labels)) + (0.3j * rng.standard_normal(len(labels)) if cplx_w else 0)
    return labels, [complex(v) for v in w]


def main():
    R = load_reference()
    WPO, Pauli = R['WPO'], terra_stub.Pauli
    out = {'note': 'outputs of the reference code itself at 11-13 qubits; see make_reference_fixtures_large.py', 'cases': []}
    for n, T, cw, seed, k in CASES:
        labels, weights = make_terms(n, T, cw, seed)
        op = WPO(paulis=[[w, Pauli(l)] for w, l in zip(weights, labels)])
        psi = make_state(n, seed)
        mean, std = op.evaluate_with_statevector(psi)
        mat = R['op_converter'].to_matrix_operator(op)
        hpsi = mat._matrix.dot(psi)
        rows = np.sort(np.random.default_rng(seed + 13).choice(1 << n, size=96, replace=False))
        res = R['NumPyEigensolver'](op, k=k).run()
        rec = {'n': n, 'seed': seed, 'labels': labels, 'weights': [cplx(w) for w in weights], 'complex_weights': cw,
               'psi_checksum': cplx(np.sum(psi * np.arange(1, (1 << n) + 1))),
               'evaluate_with_statevector': cplx(mean), 'std': std,
               'matrix_operator_evaluate': cplx(mat.evaluate_with_statevector(psi)[0]),
               'rows': [int(r) for r in rows], 'h_psi_rows': [cplx(hpsi[r]) for r in rows],
               'h_psi_norm': float(np.linalg.norm(hpsi)), 'k': k, 'eigenvalues': [cplx(v) for v in res.eigenvalues]}
        out['cases'].append(rec)
        print(n, len(labels), 'complex' if cw else 'real', 'E =', mean, 'eig', res.eigenvalues)
    path = os.path.join(HERE, 'reference_run_large.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
