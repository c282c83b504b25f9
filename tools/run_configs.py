#!/usr/bin/env python
"""Runs BASELINE.json configs 1-3 through the reference-facing entry points on cuda:0, checks them
against the oracle and prints one JSON object (profiles/configs_r1.json).  Configs 4/5 are covered
by bench.py / tools/run_config4.py / the 8-GPU bench."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co, pauli_oracle as po                      # noqa: E402
from qiskit_aqua_b200 import (NumPyEigensolver, NumPyMinimumEigensolver, Pauli, PauliSumEngine,  # noqa: E402
                              WeightedPauliOperator, max_cut, random_graph, workloads as W)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps * 1e3


def main():
    res = {}
    # ---- C1: max_cut.get_operator on a seeded 16-node graph, NumPyMinimumEigensolver ----------
    w = random_graph(16, weight_range=10, edge_prob=0.5, negative_weight=True, seed=17)
    op, shift = max_cut.get_operator(w)
    r, ms = timed(lambda: NumPyMinimumEigensolver(op).run())
    terms, _ = po.max_cut_operator(w)
    t0 = time.perf_counter()
    ref, _ = po.ref_solve(terms, k=1)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    res['C1_maxcut_16'] = {'terms': len(op.paulis), 'gpu_ms': ms, 'eigenvalue': float(np.real(r.eigenvalue)),
                           'oracle_eigenvalue': float(ref[0].real), 'oracle_cpu_ms': cpu_ms,
                           'abs_err': abs(float(np.real(r.eigenvalue)) - float(ref[0].real))}
    # ---- C2: 24-qubit VQE energy evaluation, ~2k-term 2-local sum --------------------------------
    n, x, z, y, c = W.two_local_random(24, 2000, seed=24)
    op2 = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(a), int(b), n)] for a, b, cc in zip(x, z, c)])
    psi = co.counter_state(24, 0, 1 << n)
    psi /= np.linalg.norm(psi)
    host = torch.from_numpy(psi).pin_memory()
    dev = host.cuda()
    (e_host, _), ms_host = timed(lambda: op2.evaluate_with_statevector(host))
    (e_dev, _), ms_dev = timed(lambda: op2.evaluate_with_statevector(dev))
    t0 = time.perf_counter()
    e_ref = co.expect(n, x, z, y, c, psi)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    res['C2_vqe_energy_24q'] = {'terms': len(x), 'x_groups': op2.engine.info()['n_xgroups'],
                                'ms_host_state_e2e': ms_host, 'ms_device_state': ms_dev,
                                'energy': e_host.real, 'rel_err_vs_oracle': abs(e_host - e_ref) / abs(e_ref),
                                'rel_err_dev_vs_oracle': abs(e_dev - e_ref) / abs(e_ref),
                                'oracle_c_openmp_ms': cpu_ms, 'oracle_threads': co.threads()}
    # ---- C3: 14-qubit molecular-style Hamiltonian, NumPyEigensolver k = 4 -------------------------
    n, x, z, y, c = W.molecular_style(14, seed=14)
    op3 = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(a), int(b), n)] for a, b, cc in zip(x, z, c)])
    r, ms = timed(lambda: NumPyEigensolver(op3, k=4).run(), reps=1)
    import scipy.sparse.linalg as spla
    N = 1 << n
    lin = spla.LinearOperator((N, N), dtype=np.complex128, matvec=lambda v: co.apply(n, x, z, y, c, v.astype(complex)))
    t0 = time.perf_counter()
    ev = np.sort(spla.eigsh(lin, k=6, which='SA', return_eigenvectors=False, tol=1e-13))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = np.real(r.eigenvalues)
    res['C3_molecular_14q_k4'] = {'terms': len(x), 'x_groups': op3.engine.info()['n_xgroups'], 'gpu_ms': ms,
                                  'eigenvalues': got.tolist(), 'oracle_eigsh': ev[:4].tolist(),
                                  'max_abs_err': float(np.abs(got - ev[:4]).max()),
                                  'oracle_matrix_free_eigsh_ms': cpu_ms}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
