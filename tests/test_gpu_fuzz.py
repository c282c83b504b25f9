"""GPU: randomised parity sweep over operators, sizes and planner options (seeded, deterministic)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co            # noqa: E402
from qiskit_aqua_b200 import PauliSumEngine  # noqa: E402


def _random_operator(rng, n):
    T = int(rng.integers(1, 120))
    style = rng.integers(0, 4)
    x = np.zeros(T, dtype=np.uint64)
    z = np.zeros(T, dtype=np.uint64)
    for k in range(T):
        if style == 0:                                     # dense random masks
            x[k] = rng.integers(0, 1 << n, dtype=np.uint64)
            z[k] = rng.integers(0, 1 << n, dtype=np.uint64)
        elif style == 1:                                   # low weight
            for q in rng.choice(n, size=min(n, int(rng.integers(1, 4))), replace=False):
                p = rng.integers(0, 3)
                if p < 2:
                    x[k] |= np.uint64(1 << int(q))
                if p > 0:
                    z[k] |= np.uint64(1 << int(q))
        elif style == 2:                                   # few x-masks, many z patterns
            x[k] = np.uint64(int(rng.choice([0, 1, 1 << (n - 1), (1 << n) - 1])) & ((1 << n) - 1))
            z[k] = rng.integers(0, 1 << n, dtype=np.uint64)
        else:                                              # diagonal only
            z[k] = rng.integers(0, 1 << n, dtype=np.uint64)
    ypow = np.array([bin(int(a) & int(b)).count('1') & 3 for a, b in zip(x, z)], dtype=np.uint8)
    if rng.random() < 0.3:
        ypow = rng.integers(0, 4, size=T).astype(np.uint8)            # arbitrary phase powers
    c = rng.standard_normal(T) + (1j * rng.standard_normal(T) if rng.random() < 0.5 else 0)
    return x, z, ypow, c.astype(np.complex128)


@pytest.mark.parametrize('seed', range(64))
def test_fuzz_apply_and_expect(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 15))
    x, z, ypow, c = _random_operator(rng, n)
    opts = {}
    if n >= 11:
        opts = {'tile_bits': int(rng.integers(11, min(13, n) + 1)), 'low_bits': int(rng.integers(1, 7)),
                'min_cover': int(rng.integers(1, 5)), 'late_bits': int(rng.integers(0, 3)),
                'herm_pairing': int(rng.integers(0, 2)), 'reg_bits': int(rng.integers(3, 5)),
                'cta_mode': int(rng.integers(0, 3)), 'min_bundle': int(rng.choice([2, 3, 5, 255]))}
        if rng.random() < 0.15:
            opts['kernel'] = 1
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    eng = PauliSumEngine(n, x, z, ypow, c, opts)
    ref = co.apply(n, x, z, ypow, c, psi)
    scale = max(np.abs(ref).max(), 1e-300)
    if eng.info()['n_terms'] == 0:                       # everything cancelled
        return
    got, ex = eng.apply(psi, want_expect=True)
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-10 * scale
    eref = np.vdot(psi, ref)
    tol = 1e-10 * max(abs(eref), scale)
    assert abs(ex - eref) <= tol
    assert abs(eng.expect(psi) - eref) <= tol
    assert abs(eng.expect_host(psi) - eref) <= tol
    phi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    assert abs(eng.bra_apply(phi, psi) - np.vdot(phi, ref)) <= 1e-10 * max(1.0, np.abs(phi).max() * scale * (1 << n))


@pytest.mark.parametrize('world,n,seed', [(2, 12, 1), (4, 13, 2), (8, 14, 3), (4, 9, 4)])
def test_sharded_parts_emulated_on_one_gpu(world, n, seed):
    """The N > 1 kernel path without N GPUs: every rank's shard lives on this device and the
    'exchange' is a pointer to the partner shard (psum_apply_part).  Checks the part split, the
    rank-sign folding, remote-part passes reading a different source than the bra, and the
    expectation semantics (with y: <bra|y> after the last part; without y: per-part contributions)."""
    rng = np.random.default_rng(seed)
    x, z, ypow, c = _random_operator(rng, n)
    x[0] = np.uint64((1 << n) - 1)                           # make sure global x-parts occur
    ypow = np.array([bin(int(a) & int(b)).count('1') & 3 for a, b in zip(x, z)], dtype=np.uint8)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    ref = co.apply(n, x, z, ypow, c, psi)
    scale = np.abs(ref).max()
    N_loc = (1 << n) // world
    full = torch.from_numpy(psi).cuda()
    shards = [full[r * N_loc:(r + 1) * N_loc].contiguous() for r in range(world)]
    e_with_y = 0.0
    e_contrib = 0.0
    for r in range(world):
        eng = PauliSumEngine(n, x, z, ypow, c, {'reg_bits': 3 + (seed + r) % 2})
        eng.shard(r, world)
        y = torch.zeros(N_loc, dtype=torch.complex128, device='cuda')
        parts = [g for g in range(world) if len(eng.shard_terms(g)[0])]
        last = 0.0
        for i, g in enumerate(parts):
            last = eng.apply_part(g, shards[r ^ g], shards[r], y, accumulate=True, want_expect=True)
            e_contrib += eng.apply_part(g, shards[r ^ g], shards[r], None, accumulate=False, want_expect=True)
        e_with_y += last
        got = y.cpu().numpy()
        assert np.abs(got - ref[r * N_loc:(r + 1) * N_loc]).max() <= 1e-10 * scale
    eref = np.vdot(psi, ref)
    assert abs(e_with_y - eref) <= 1e-10 * max(abs(eref), scale)
    assert abs(e_contrib - eref) <= 1e-10 * max(abs(eref), scale)
