"""Reference-run fixtures at 11-13 qubits (tests/golden/make_reference_fixtures_large.py ->
tests/golden/reference_run_large.json): the reference's own CSR path (WeightedPauliOperator ->
op_converter.to_matrix_operator -> csr.dot -> np.vdot) and its NumPyEigensolver (scipy eigs, real and
complex weights) executed in the build container, at sizes where the product runs the TILED kernel
(k_pass) -- reference_run.json stops at 6 qubits, the streaming kernel.
CPU part: both oracles reproduce the recorded values.  GPU part: so do the kernels, through the
reference-facing entry points (Lanczos for real weights, psum_arnoldi for complex ones)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import pauli_oracle as po
from qiskit_aqua_b200 import Pauli, WeightedPauliOperator

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'reference_run_large.json')) as _f:
    CASES = json.load(_f)['cases']


def _c(pair):
    return complex(pair[0], pair[1])


def _state(rec):
    """The generator's state, regenerated from its seed (checked against the recorded checksum)."""
    n = rec['n']
    rng = np.random.default_rng(rec['seed'])
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    chk = np.sum(psi * np.arange(1, (1 << n) + 1))
    assert abs(chk - _c(rec['psi_checksum'])) <= 1e-9 * abs(chk)
    return psi


def _terms(rec):
    return [(_c(w), lab) for w, lab in zip(rec['weights'], rec['labels'])]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_oracles_vs_large_reference_run(case):
    rec = CASES[case]
    n = rec['n']
    psi = _state(rec)
    terms = po.simplify(_terms(rec))
    x, z, y, c = po.pack_terms(terms)
    rows = np.array(rec['rows'])
    want_rows = np.array([_c(v) for v in rec['h_psi_rows']])
    ref_e = _c(rec['evaluate_with_statevector'])
    assert abs(_c(rec['matrix_operator_evaluate']) - ref_e) < 1e-12
    for hpsi in (po.mf_apply(n, x, z, y, c, psi), co.apply(n, x, z, y, c, psi)):
        np.testing.assert_allclose(hpsi[rows], want_rows, rtol=0, atol=1e-12)
        assert abs(np.linalg.norm(hpsi) - rec['h_psi_norm']) < 1e-11
        assert abs(np.vdot(psi, hpsi) - ref_e) < 1e-12
    val, _ = po.ref_solve(terms, k=rec['k'])
    ref = np.array([_c(v) for v in rec['eigenvalues']])
    val = np.asarray(val, dtype=complex)
    val = val[np.lexsort((val.imag, val.real))]
    np.testing.assert_allclose(val, ref, rtol=0, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize('case', range(len(CASES)))
def test_tiled_kernels_vs_large_reference_run(case):
    pytest.importorskip('torch')
    rec = CASES[case]
    psi = _state(rec)
    op = WeightedPauliOperator(paulis=[[w, Pauli(lab)] for w, lab in _terms(rec)])
    assert op.engine.info()['n_passes'] >= 1                      # the tiled kernel serves this size
    mean, std = op.evaluate_with_statevector(psi)
    ref_e = _c(rec['evaluate_with_statevector'])
    assert std == 0.0 and abs(mean - ref_e) <= 1e-10 * max(1.0, abs(ref_e))
    y = op.engine.apply(psi).cpu().numpy()
    want_rows = np.array([_c(v) for v in rec['h_psi_rows']])
    assert np.abs(y[np.array(rec['rows'])] - want_rows).max() <= 1e-10 * max(1.0, np.abs(want_rows).max())
    assert abs(np.linalg.norm(y) - rec['h_psi_norm']) <= 1e-10 * rec['h_psi_norm']


@pytest.mark.gpu
@pytest.mark.parametrize('case', range(len(CASES)))
def test_product_eigensolver_vs_large_reference_run(case):
    pytest.importorskip('torch')
    from qiskit_aqua_b200 import NumPyEigensolver
    rec = CASES[case]
    op = WeightedPauliOperator(paulis=[[w, Pauli(lab)] for w, lab in _terms(rec)])
    res = NumPyEigensolver(op, k=rec['k']).run()
    ref = np.array([_c(v) for v in rec['eigenvalues']])
    got = np.asarray(res.eigenvalues, dtype=complex)
    assert len(got) == len(ref)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-8 * max(1.0, np.abs(ref).max()))
