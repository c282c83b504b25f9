"""GPU: Krylov time evolution exp(-i t H) psi against scipy expm (the reference's exact branch,
legacy/matrix_operator.py:354-355) and size-independent properties."""
import numpy as np
import pytest
import scipy.linalg as scila

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import pauli_oracle as po        # noqa: E402
from qiskit_aqua_b200 import Pauli, PauliSumEngine, WeightedPauliOperator, workloads as W  # noqa: E402


@pytest.mark.parametrize('n,t', [(2, 0.7), (6, 1.3), (9, -2.5)])
def test_evolve_matches_expm(n, t):
    rng = np.random.default_rng(n)
    labels = set()
    while len(labels) < min(30, 4 ** n - 1):
        labels.add(''.join(rng.choice(list('IXYZ')) for _ in range(n)))
    labels = sorted(labels)
    weights = rng.standard_normal(len(labels))
    op = WeightedPauliOperator.from_list([Pauli(l) for l in labels], list(weights))
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    H = po.ref_to_spmatrix(list(zip(weights, labels)))
    ref = scila.expm(-1j * t * H.toarray()) @ psi
    got = op.evolve(psi, evo_time=t)
    assert np.abs(got - ref).max() < 1e-8
    with pytest.raises(ValueError):
        op.evolve(psi, evo_time=t, num_time_slices=-1)


def test_evolve_is_unitary_and_reversible_20q():
    n, x, z, y, c = W.two_local_random(20, 400, seed=5)
    eng = PauliSumEngine(n, x, z, y, c)
    rng = np.random.default_rng(1)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    psi_d = torch.from_numpy(psi).cuda()
    e0 = eng.expect(psi_d)
    fwd, info = eng.evolve(psi_d, 0.05, tol=1e-11)
    assert info['steps'] >= 1 and info['matvecs'] >= 2
    assert abs(torch.linalg.vector_norm(fwd).item() - 1.0) < 1e-9      # unitary
    assert abs(eng.expect(fwd) - e0) < 1e-8 * max(1.0, abs(e0))         # energy conserved
    back, _ = eng.evolve(fwd, -0.05, tol=1e-11)
    assert torch.linalg.vector_norm(back - psi_d).item() < 1e-8          # reversible
