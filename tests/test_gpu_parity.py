"""GPU parity tests proper: the CUDA path, called through the C ABI (libpsum.so via ctypes),
against the CPU oracle on the same seeded inputs.  Tolerance: 1e-10 relative (north_star)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co            # noqa: E402
from oracle import pauli_oracle as po        # noqa: E402
from qiskit_aqua_b200 import PauliSumEngine, fill_state, workloads as W  # noqa: E402
from qiskit_aqua_b200.engine import vec_dot  # noqa: E402

RTOL = 1e-10


def _state(n, seed):
    rng = np.random.default_rng(seed)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return psi / np.linalg.norm(psi)


def _check(n, x, z, y, c, options=None, seed=0):
    psi = _state(n, seed)
    eng = PauliSumEngine(n, x, z, y, c, options)
    ref = co.apply(n, x, z, y, c, psi)
    got, ex = eng.apply(psi, want_expect=True)
    got = got.cpu().numpy()
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.abs(got - ref).max() <= RTOL * scale
    eref = np.vdot(psi, ref)
    assert abs(ex - eref) <= RTOL * max(abs(eref), scale)
    e2 = eng.expect(psi)
    assert abs(e2 - eref) <= RTOL * max(abs(eref), scale)
    return eng


@pytest.mark.parametrize('n', [1, 2, 3, 5, 8, 10, 11, 12, 13, 14])
def test_random_complex_operator(n):
    T = min(4 ** n, 150)
    _, x, z, y, c = W.random_dense_paulis(n, T, seed=100 + n, complex_coeffs=True)
    _check(n, x, z, y, c, seed=n)


@pytest.mark.parametrize('reg_bits', [3, 4])
@pytest.mark.parametrize('tile_bits,low_bits', [(11, 4), (12, 4), (13, 4), (12, 1), (12, 6), (13, 3)])
def test_two_local_tiled(tile_bits, low_bits, reg_bits):
    n, x, z, y, c = W.two_local_random(16, 500, seed=16)
    eng = _check(n, x, z, y, c, {'tile_bits': tile_bits, 'low_bits': low_bits, 'reg_bits': reg_bits})
    info = eng.info()
    assert info['n_passes'] >= 1 and info['tile_bits'] == tile_bits
    assert info['n_bundles'] < info['n_group_records']        # partner loads are shared


@pytest.mark.parametrize('term_cache', [0, 1])
@pytest.mark.parametrize('cta_mode', [0, 1])
def test_term_cache_on_and_off(term_cache, cta_mode):
    """One-chunk real-only passes keep their component terms in shared memory (option term_cache, default
    on); both settings give the oracle's H.psi and <H>, as one big CTA per SM and as two CTAs per SM."""
    n, x, z, y, c = W.ising_heisenberg(18, n_triples=3000, seed=30)
    _check(n, x, z, y, c, {'term_cache': term_cache, 'cta_mode': cta_mode})


@pytest.mark.parametrize('cta_mode', [0, 1, 2])
@pytest.mark.parametrize('reg_bits', [3, 4])
@pytest.mark.parametrize('kind', ['c4', 'c5', 'molecular'])
def test_benchmark_families_both_register_widths(kind, reg_bits, cta_mode):
    if kind == 'c4':
        n, x, z, y, c = W.ising_heisenberg(18, n_triples=3000, seed=30)
    elif kind == 'c5':
        n, x, z, y, c = W.random_weighted_paulis(18, 3000, 4, seed=34)
    else:
        n, x, z, y, c = W.molecular_style(14)
    eng = _check(n, x, z, y, c, {'reg_bits': reg_bits, 'cta_mode': cta_mode})
    if kind == 'c4':
        assert eng.info()['n_quad_groups'] >= n          # separable-table path is exercised
    psi = _state(n, 5)
    ref = np.vdot(psi, co.apply(n, x, z, y, c, psi))
    assert abs(eng.expect_host(psi) - ref) <= RTOL * max(1.0, abs(ref))


def test_streaming_kernel_forced():
    n, x, z, y, c = W.two_local_random(14, 300, seed=3)
    eng = _check(n, x, z, y, c, {'kernel': 1})
    assert eng.info()['n_passes'] == 0


def test_remote_groups_and_wide_masks():
    # x-masks wider than a tile cannot be served from shared memory -> global partner reads
    n = 15
    rng = np.random.default_rng(5)
    xs = [(1 << n) - 1, ((1 << n) - 1) ^ 0b1010, 0b111111111111111 & 0x5555, 1 << 14, 0]
    x = np.array(xs, dtype=np.uint64)
    z = rng.integers(0, 1 << n, size=len(xs), dtype=np.uint64)
    y = np.array([bin(int(a) & int(b)).count('1') & 3 for a, b in zip(x, z)], dtype=np.uint8)
    c = rng.standard_normal(len(xs)) + 1j * rng.standard_normal(len(xs))
    eng = _check(n, x, z, y, c, {'tile_bits': 12})
    assert eng.info()['n_remote_groups'] >= 2


def test_big_diagonal_group_is_split():
    # one x-group with thousands of Z patterns: exercises sub-group / chunk splitting
    n = 13
    rng = np.random.default_rng(9)
    z = np.unique(rng.integers(0, 1 << n, size=4000, dtype=np.uint64))
    x = np.zeros_like(z)
    y = np.zeros(len(z), dtype=np.uint8)
    c = rng.standard_normal(len(z)).astype(np.complex128)
    eng = _check(n, x, z, y, c)
    assert eng.info()['is_diagonal'] == 1
    d = eng.diag().cpu().numpy()
    ref = po.mf_diag(n, z, c).real
    assert np.abs(d - ref).max() <= RTOL * np.abs(ref).max()
    vals, idx = eng.diag_kmin(5)
    order = np.lexsort((np.arange(1 << n), ref))[:5]
    np.testing.assert_array_equal(idx, order.astype(np.uint64))
    np.testing.assert_allclose(vals, ref[order], rtol=0, atol=RTOL * np.abs(ref).max())


def test_many_classes_imaginary_parts():
    # Y-heavy operator: imaginary c' and every (zr) class populated
    n = 12
    rng = np.random.default_rng(12)
    labels = []
    while len(labels) < 300:
        lab = ''.join(rng.choice(list('IXYZ'), p=[.4, .1, .4, .1]) for _ in range(n))
        if lab not in labels:
            labels.append(lab)
    coeffs = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    x, z, y = __import__('qiskit_aqua_b200').pack_masks(labels)
    _check(n, x, z, y, coeffs)


def test_duplicates_merged_and_zero_dropped():
    # simplify semantics (legacy/weighted_pauli_operator.py:341-400)
    n = 4
    labels = ['XXII', 'XXII', 'ZIZI', 'ZIZI', 'IIII']
    coeffs = [0.5, 0.25, 1.0, -1.0, 2.0]
    eng = PauliSumEngine.from_labels(labels, coeffs)
    assert eng.info()['n_terms'] == 2
    psi = _state(n, 1)
    terms = po.simplify(list(zip(coeffs, labels)))
    ref = po.ref_evaluate_with_statevector(terms, psi)[0]
    assert abs(eng.expect(psi) - ref) < 1e-12


def test_bra_apply_and_hermiticity():
    n, x, z, y, c = W.two_local_random(13, 200, seed=2)
    eng = PauliSumEngine(n, x, z, y, c)
    psi, phi = _state(n, 1), _state(n, 2)
    ref = np.vdot(phi, co.apply(n, x, z, y, c, psi))
    got = eng.bra_apply(phi, psi)
    assert abs(got - ref) <= RTOL * max(1.0, abs(ref))
    assert abs(eng.bra_apply(psi, phi) - np.conj(got)) <= 1e-12


def test_expect_batch():
    n, x, z, y, c = W.two_local_random(12, 100, seed=4)
    eng = PauliSumEngine(n, x, z, y, c)
    psis = np.stack([_state(n, s) for s in range(3)])
    got = eng.expect_batch(torch.from_numpy(psis))
    ref = [co.expect(n, x, z, y, c, p) for p in psis]
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-11)


def test_fill_state_twin_is_bit_exact():
    a = fill_state(1 << 12, seed=77, first_index=12345).cpu().numpy()
    b = co.counter_state(77, 12345, 1 << 12)
    np.testing.assert_array_equal(a, b)


def test_empty_operator_raises():
    eng = PauliSumEngine(3, np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint8),
                         np.zeros(0, np.complex128))
    with pytest.raises(Exception) as ei:
        eng.expect(np.ones(8, dtype=complex))
    assert 'empty' in str(ei.value).lower()


def test_config2_24q_energy():
    # BASELINE config 2: 24-qubit seeded state, ~2k-term 2-local sum; oracle = C restatement
    n, x, z, y, c = W.two_local_random(24, 2000, seed=24)
    eng = PauliSumEngine(n, x, z, y, c)
    N = 1 << n
    psi = fill_state(N, seed=24)
    nrm = np.sqrt(vec_dot(psi, psi).real)
    yv, ex = eng.apply(psi, want_expect=True)
    rng = np.random.default_rng(0)
    rows = rng.integers(0, N, size=256, dtype=np.uint64)
    ref = co.apply_rows_counter(x, z, y, c, 24, 1.0, rows)
    got = yv[torch.from_numpy(rows.astype(np.int64)).cuda()].cpu().numpy()
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    # energy against the full C oracle on the host copy
    eref = co.expect(n, x, z, y, c, psi.cpu().numpy())
    assert abs(ex - eref) <= RTOL * abs(eref)
    assert abs(eng.expect(psi) - eref) <= RTOL * abs(eref)
    assert nrm > 0


@pytest.mark.parametrize('n', [28, 30])
def test_full_size_sampled_rows_and_properties(n):
    """BASELINE full size (30 q, 10k terms): sampled rows against the oracle with the
    counter-based state, plus size-independent properties (Hermiticity, linearity)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 3.2 * 16 * (1 << n):
        pytest.skip('not enough device memory for %d qubits' % n)
    _, x, z, y, c = W.ising_heisenberg(n, n_triples=8665, seed=30)
    eng = PauliSumEngine(n, x, z, y, c)
    N = 1 << n
    psi = fill_state(N, seed=30)
    yv, ex = eng.apply(psi, want_expect=True)
    rng = np.random.default_rng(1)
    rows = np.concatenate([rng.integers(0, N, size=192, dtype=np.uint64),
                           np.array([0, 1, N - 1, N // 2, (N // 3) | 1], dtype=np.uint64)])
    ref = co.apply_rows_counter(x, z, y, c, 30, 1.0, rows)
    got = yv[torch.from_numpy(rows.astype(np.int64)).cuda()].cpu().numpy()
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    # <psi|H|psi> is real for a Hermitian H and equals <psi|y>
    assert abs(ex.imag) <= 1e-10 * abs(ex.real)
    assert abs(vec_dot(psi, yv) - ex) <= RTOL * abs(ex)
    assert abs(eng.expect(psi) - ex) <= RTOL * abs(ex)
    # linearity: H(2 psi) = 2 H psi  -> energy scales by 4
    psi *= 2.0
    assert abs(eng.expect(psi) - 4 * ex) <= RTOL * abs(4 * ex)


@pytest.mark.parametrize('n,kind', [(12, 'numpy'), (20, 'numpy'), (21, 'pinned'), (22, 'complex')])
def test_expect_host_pipelined(n, kind):
    """psum_expect_host: chunked H2D overlapped with the chunk-local passes must give the same
    <psi|H|psi> as the resident-state kernels and the oracle."""
    if kind == 'complex':
        _, x, z, y, c = W.random_weighted_paulis(n, 300, 3, seed=n)
        c = c * (1.0 + 0.25j)                      # non-Hermitian: generic (unpaired) kernels
    else:
        _, x, z, y, c = W.two_local_random(n, 400, seed=n)
    eng = PauliSumEngine(n, x, z, y, c)
    psi = _state(n, n)
    ref = co.expect(n, x, z, y, c, psi)
    host = torch.from_numpy(psi).pin_memory() if kind == 'pinned' else psi
    got = eng.expect_host(host)
    assert abs(got - ref) <= RTOL * max(abs(ref), 1e-3)
    assert abs(eng.expect(psi) - ref) <= RTOL * max(abs(ref), 1e-3)
    got2 = eng.expect_host(host)                   # second call reuses the plan and the streams
    assert abs(got2 - got) <= 1e-13 * max(abs(got), 1e-3)


@pytest.mark.parametrize('kind,n,opts', [('c4', 16, {}), ('c4', 16, {'reg_bits': 4}), ('c5', 15, {}), ('c5', 15, {'cta_mode': 1}),
                                         ('molecular', 14, {}), ('small', 8, {}), ('c4', 17, {'tile_bits': 13})])
def test_batched_expectation_shares_the_staged_plan(kind, n, opts):
    """psum_expect_batch: one launch per pass for up to 32 states (the fold / table build of a tile is
    shared by the batch) against the per-state expectation and the oracle; batches above 32 states and
    passes with several chunks take the per-state launches."""
    if kind == 'c4':
        n, x, z, y, c = W.ising_heisenberg(n, n_triples=2500, seed=3)
    elif kind == 'c5':
        n, x, z, y, c = W.random_weighted_paulis(n, 2500, 4, seed=5)
    elif kind == 'molecular':
        n, x, z, y, c = W.molecular_style(n)
    else:
        n, x, z, y, c = W.two_local_random(n, 40, seed=1)
    eng = PauliSumEngine(n, x, z, y, c, opts)
    rng = np.random.default_rng(n)
    for batch in (1, 5, 40):
        psis = rng.standard_normal((batch, 1 << n)) + 1j * rng.standard_normal((batch, 1 << n))
        psis /= np.linalg.norm(psis, axis=1)[:, None]
        dev = torch.from_numpy(psis).cuda()
        got = eng.expect_batch(dev)
        one = np.array([eng.expect(dev[b]) for b in range(batch)])
        ref = np.array([np.vdot(psis[b], co.apply(n, x, z, y, c, psis[b])) for b in range(min(batch, 3))])
        scale = max(1.0, np.abs(one).max())
        assert np.abs(got - one).max() <= 1e-12 * scale
        assert np.abs(got[:len(ref)] - ref).max() <= RTOL * max(1.0, np.abs(ref).max())
