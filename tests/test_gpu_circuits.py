"""GPU: ansatz -> psi on the device against a plain numpy gate simulator, and the VQE-style
energy evaluation against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import pauli_oracle as po        # noqa: E402
from qiskit_aqua_b200 import (CircuitStateFn, EfficientSU2, Pauli, RealAmplitudes, StateFn,  # noqa: E402
                              StatevectorCircuit, WeightedPauliOperator, energy_evaluation)
from qiskit_aqua_b200.circuits import _u    # noqa: E402


def _np_run(circ, params):
    n = circ.num_qubits
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    for name, qs, val in circ.gates:
        if name == 'cx':
            c, t = qs
            sel = idx[((idx >> c) & 1 == 1) & ((idx >> t) & 1 == 0)]
            psi[sel], psi[sel | (1 << t)] = psi[sel | (1 << t)].copy(), psi[sel].copy()
        elif name == 'cz':
            c, t = qs
            psi[((idx >> c) & 1 == 1) & ((idx >> t) & 1 == 1)] *= -1
        else:
            theta = 0.0 if val is None else (params[val[1]] if isinstance(val, tuple) else val)
            u = np.asarray(_u(name, theta), dtype=complex)
            q = qs[0]
            i0 = idx[(idx >> q) & 1 == 0]
            a, b = psi[i0].copy(), psi[i0 | (1 << q)].copy()
            psi[i0] = u[0, 0] * a + u[0, 1] * b
            psi[i0 | (1 << q)] = u[1, 0] * a + u[1, 1] * b
    return psi


@pytest.mark.parametrize('n', [1, 2, 5, 9])
def test_gates_match_numpy(n):
    rng = np.random.default_rng(n)
    circ = StatevectorCircuit(n)
    for _ in range(40):
        g = rng.choice(['ry', 'rx', 'rz', 'h', 'x', 'cx', 'cz'])
        if g in ('cx', 'cz'):
            if n < 2:
                continue
            c, t = rng.choice(n, size=2, replace=False)
            getattr(circ, g)(int(c), int(t))
        elif g in ('h', 'x'):
            getattr(circ, g)(int(rng.integers(n)))
        else:
            getattr(circ, g)(float(rng.standard_normal()), int(rng.integers(n)))
    got = circ.run().cpu().numpy()
    np.testing.assert_allclose(got, _np_run(circ, []), atol=1e-13)
    np.testing.assert_allclose(circ.run_unfused().cpu().numpy(), got, atol=1e-13)


def test_ghz_and_ansatz_energy():
    ghz = StatevectorCircuit(4).h(0).cx(0, 1).cx(1, 2).cx(2, 3)
    psi = ghz.run().cpu().numpy()
    ref = np.zeros(16, dtype=complex)
    ref[0] = ref[15] = 1 / np.sqrt(2)
    np.testing.assert_allclose(psi, ref, atol=1e-14)
    # VQE-style energy evaluation: H2 operator, RealAmplitudes / EfficientSU2 ansatz
    h2 = [(-1.052373245772859, 'II'), (0.39793742484318045, 'ZI'), (-0.39793742484318045, 'IZ'),
          (-0.01128010425623538, 'ZZ'), (0.18093119978423156, 'XX')]
    op = WeightedPauliOperator(paulis=[[w, Pauli(l)] for w, l in h2])
    rng = np.random.default_rng(0)
    for ansatz in (RealAmplitudes(2, reps=2), EfficientSU2(2, reps=1, entanglement='linear')):
        theta = rng.standard_normal(2 * ansatz.num_parameters)
        means = energy_evaluation(op, ansatz, theta)            # two grouped parameter sets
        assert means.shape == (2,)
        for k in range(2):
            p = theta[k * ansatz.num_parameters:(k + 1) * ansatz.num_parameters]
            ref_e = po.ref_evaluate_with_statevector(h2, _np_run(ansatz, p))[0].real
            assert abs(means[k] - ref_e) < 1e-12
        val = (~StateFn(op) @ CircuitStateFn(ansatz, theta[:ansatz.num_parameters])).eval()
        assert abs(val.real - means[0]) < 1e-12


def test_vqe_h2_reaches_reference_energy():
    # test/aqua/test_vqe.py:38-60: H2 qubit operator, RY-RZ ansatz, SLSQP -> -1.85727503
    from qiskit_aqua_b200 import VQE
    h2 = [(-1.052373245772859, 'II'), (0.39793742484318045, 'ZI'), (-0.39793742484318045, 'IZ'),
          (-0.01128010425623538, 'ZZ'), (0.18093119978423156, 'XX')]
    op = WeightedPauliOperator(paulis=[[w, Pauli(l)] for w, l in h2])
    best = min(VQE(op, EfficientSU2(2, reps=2, entanglement='linear'), optimizer='SLSQP', seed=s).run()['optimal_value']
               for s in range(3))
    assert abs(best - (-1.85727503)) < 1e-5


@pytest.mark.parametrize('n,ent,maxq', [(14, 'linear', 12), (15, 'full', 12), (16, 'circular', 6), (13, 'linear', 5)])
def test_fused_blocks_equal_gate_by_gate(n, ent, maxq):
    """Gate fusion (one sweep per block of gates whose qubits fit a shared-memory tile) against the
    gate-by-gate path and the numpy simulator; the scheduler keeps program order wherever gates share
    a qubit."""
    from qiskit_aqua_b200.circuits import fuse_schedule
    rng = np.random.default_rng(n)
    circ = EfficientSU2(n, reps=2, entanglement=ent)
    for _ in range(6):                                   # a few gates the ansatz family lacks
        c, t = rng.choice(n, size=2, replace=False)
        circ.cz(int(c), int(t)).h(int(rng.integers(n))).rx(float(rng.standard_normal()), int(rng.integers(n)))
    theta = rng.uniform(-np.pi, np.pi, circ.num_parameters)
    fused = circ.run(theta, max_block_qubits=maxq).cpu().numpy()
    assert circ.last_block_count < len(circ.gates) / 3
    np.testing.assert_allclose(fused, circ.run_unfused(theta).cpu().numpy(), atol=1e-13)
    np.testing.assert_allclose(fused, _np_run(circ, theta), atol=1e-13)
    blocks = fuse_schedule(n, circ._bound_gates(theta), maxq)
    assert all(len(f) <= max(maxq, 5) and set(q for g in b for q in g[1]) <= set(f) for f, b in blocks)


def test_fused_state_preparation_20_qubits():
    circ = EfficientSU2(20, reps=3, entanglement='linear')
    theta = np.random.default_rng(1).uniform(-np.pi, np.pi, circ.num_parameters)
    a = circ.run(theta)
    b = circ.run_unfused(theta)
    assert circ.last_block_count <= 12
    assert float(torch.linalg.vector_norm(a - b).item()) < 1e-12
    assert abs(float(torch.linalg.vector_norm(a).item()) - 1.0) < 1e-12
