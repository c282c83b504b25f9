"""CPU verification of the pass planner: the arrays `k_pass` reads (exported with
psum_plan_export) are executed by the numpy interpreter in tests/plan_interpreter.py and compared
with the oracle -- tile addressing, fold lists, bundles (one partner set, a register permutation per
member), class entries, inline fast slots, chunking, sub-groups, remote bundles, Hermitian pairing
flags, late-bit plans and sharded parts; for 8 and 16 amplitudes per thread (reg_bits 3 and 4)."""
import numpy as np
import pytest

from oracle import pauli_oracle as po
from qiskit_aqua_b200.engine import PauliSumEngine

from tests import plan_interpreter as pi


def random_terms(n, T, rng, max_weight=4, complex_weights=True, x_pool=None, dense=0, no_y=False):
    """Packed random Pauli terms; x_pool limits the distinct x-masks (many z-patterns per group)."""
    x = np.zeros(T, dtype=np.uint64)
    z = np.zeros(T, dtype=np.uint64)
    for k in range(T):
        if k < dense:
            qs = rng.choice(n, size=min(n, 10), replace=False)
        else:
            qs = rng.choice(n, size=rng.integers(1, max_weight + 1), replace=False)
        kinds = rng.integers(1, 4, size=len(qs))          # 1 X, 2 Y, 3 Z
        if no_y:
            kinds[kinds == 2] = 1
        for q, kind in zip(qs, kinds):
            if kind in (1, 2):
                x[k] |= np.uint64(1) << np.uint64(q)
            if kind in (2, 3):
                z[k] |= np.uint64(1) << np.uint64(q)
    if x_pool is not None:
        pool = x[:x_pool].copy()
        x = pool[rng.integers(0, x_pool, size=T)]
        z = rng.integers(0, 1 << n, size=T, dtype=np.uint64)
    ypow = np.array([bin(int(a) & int(b)).count('1') & 3 for a, b in zip(x, z)], dtype=np.uint8)
    c = rng.standard_normal(T) + (1j * rng.standard_normal(T) if complex_weights else 0)
    return x, z, ypow, c.astype(np.complex128)


def random_state(n, rng):
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return psi / np.linalg.norm(psi)


CASES = [
    # n, T, options, generator kwargs
    (11, 300, {'tile_bits': 11}, {}),
    (12, 800, {'tile_bits': 11}, {}),
    (13, 500, {'tile_bits': 11, 'low_bits': 4}, {'max_weight': 5}),
    (13, 400, {'tile_bits': 12}, {'complex_weights': False}),
    (12, 3000, {'tile_bits': 12}, {}),                                   # > 1024 patterns: several chunks
    (12, 1500, {'tile_bits': 11}, {'x_pool': 6}),                        # many classes per group: sub-groups
    (14, 200, {'tile_bits': 11}, {'dense': 12}),                         # x-masks no pass can hold: remote groups
    (13, 600, {'tile_bits': 11, 'late_bits': 2}, {}),                    # host-state plan (early passes chunk-local)
    (14, 600, {'tile_bits': 11, 'late_bits': 2, 'late_split': 1}, {}),   # host-state plan, one late qubit per late pass
    (13, 300, {'tile_bits': 13}, {}),
    (12, 600, {'tile_bits': 11}, {'no_y': True, 'complex_weights': False}),   # real-only passes (IM = false kernels)
    (11, 300, {'tile_bits': 11, 'reg_bits': 4}, {}),
    (13, 500, {'tile_bits': 12, 'reg_bits': 4}, {'max_weight': 5}),
    (12, 3000, {'tile_bits': 12, 'reg_bits': 4}, {}),
    (12, 1500, {'tile_bits': 11, 'reg_bits': 4}, {'x_pool': 6}),
    (14, 200, {'tile_bits': 11, 'reg_bits': 4}, {'dense': 12}),
    (13, 300, {'tile_bits': 13, 'reg_bits': 4}, {}),
    (12, 600, {'tile_bits': 11, 'reg_bits': 4}, {'no_y': True, 'complex_weights': False}),
    (12, 800, {'tile_bits': 11, 'min_bundle': 4}, {}),
    (13, 500, {'tile_bits': 12, 'reg_bits': 4, 'min_bundle': 255}, {'max_weight': 5}),
]


@pytest.mark.parametrize('n,T,opts,kw', CASES, ids=['%dq-%dt-%s' % (c[0], c[1], '-'.join('%s%s' % kv for kv in c[2].items())) for c in CASES])
def test_exported_plan_reproduces_oracle(n, T, opts, kw):
    rng = np.random.default_rng(1000 * n + T)
    x, z, ypow, c = random_terms(n, T, rng, **kw)
    eng = PauliSumEngine(n, x, z, ypow, c, options=opts)
    info = eng.info()
    passes = pi.export_part(eng, 0)
    assert len(passes) == info['n_passes'] > 0
    assert sum(len(ps['groups']) for ps in passes) == info['n_group_records']
    singles = sum(int(ch['ns']) for ps in passes for ch in ps['chunks'])
    assert sum(len(ps['bundles']) for ps in passes) + singles == info['n_bundles'] <= info['n_group_records']
    assert all(ps['rb'] == opts.get('reg_bits', 3) for ps in passes)
    assert sum(len(ps['pat_zt']) for ps in passes) == info['n_patterns']
    assert sum(len(ps['term_val']) for ps in passes) == info['n_component_terms']     # every term exactly once
    if kw.get('dense'):
        assert info['n_remote_groups'] > 0
    if kw.get('no_y'):
        assert not any(ps['has_im'] for ps in passes)
    if kw.get('x_pool'):
        assert info['n_group_records'] > info['n_xgroups']          # groups split into sub-groups
    if T >= 3000:
        assert max(len(ps['chunks']) for ps in passes) > 2         # staged in several chunks
    if 'late_bits' in opts:      # early passes never combine amplitudes of different chunks
        cbits = n - opts['late_bits']
        early = [ps for ps in passes if ps['hi_bit'] < cbits]
        assert early and all(ps['hi_bit'] < cbits for ps in passes[:len(early)])
    psi = random_state(n, rng)
    y = np.zeros(1 << n, dtype=np.complex128)
    e = pi.run_part(passes, psi, y)
    want = po.mf_apply(n, x, z, ypow, c, psi)
    scale = np.linalg.norm(want)
    assert np.linalg.norm(y - want) <= 1e-12 * scale
    assert abs(e - np.vdot(psi, want)) <= 1e-12 * scale


@pytest.mark.parametrize('n,T,opts', [(11, 400, {'tile_bits': 11}), (13, 700, {'tile_bits': 11}), (12, 2500, {'tile_bits': 12}),
                                      (13, 700, {'tile_bits': 12, 'reg_bits': 4})])
def test_hermitian_pairing_flags(n, T, opts):
    """Expectation with the pairing the HERM kernels use: groups flagged with a warp-uniform top
    x bit are evaluated on half of the amplitudes with doubled patterns; equals Re <psi|H|psi>."""
    rng = np.random.default_rng(77 * n + T)
    x, z, ypow, c = random_terms(n, T, rng, complex_weights=False)
    eng = PauliSumEngine(n, x, z, ypow, c, options=opts)
    assert eng.info()['is_hermitian'] == 1
    passes = pi.export_part(eng, 0)
    flagged = sum(int(np.count_nonzero(ps['groups']['meta'] >> 28)) for ps in passes)
    assert flagged > 0
    for ps in passes:       # bit 15 of a stored pattern <=> its group is paired
        for ch in ps['chunks']:
            for G in ps['groups'][int(ch['g0']):int(ch['g0']) + int(ch['ng'])]:
                first = int(ch['p0']) + (int(G['meta']) & 0xffff)
                count = sum(int(e) & 0x7ff for e in G['ent'][:(int(G['meta']) >> 16) & 7])
                marked = ps['pat_zt'][first:first + count] >> 15
                assert np.all(marked == (1 if int(G['meta']) >> 28 else 0))
    psi = random_state(n, rng)
    want = np.vdot(psi, po.mf_apply(n, x, z, ypow, c, psi))
    got = pi.run_part(passes, psi, None, herm=True)
    assert got.imag == 0.0 and abs(got.real - want.real) <= 1e-12 * max(1.0, abs(want))


@pytest.mark.parametrize('rb', [3, 4])
@pytest.mark.parametrize('world', [2, 4])
def test_sharded_parts_reproduce_oracle(world, rb):
    """Every rank's parts (global x-part g: partner shard r ^ g, rank sign folded into the weights)
    executed by the interpreter give that rank's slice of H.psi."""
    n, T = 13, 500
    rng = np.random.default_rng(5 + world)
    x, z, ypow, c = random_terms(n, T, rng)
    psi = random_state(n, rng)
    want = po.mf_apply(n, x, z, ypow, c, psi)
    n_loc = n - (world.bit_length() - 1)
    shards = psi.reshape(world, 1 << n_loc)
    seen_parts = set()
    for rank in range(world):
        eng = PauliSumEngine(n, x, z, ypow, c, options={'tile_bits': 11, 'reg_bits': rb})
        eng.shard(rank, world)
        y = np.zeros(1 << n_loc, dtype=np.complex128)
        for g in range(world):
            passes = pi.export_part(eng, g)
            if passes:
                seen_parts.add(g)
                pi.run_part(passes, shards[rank ^ g], y)
        assert np.linalg.norm(y - want.reshape(world, -1)[rank]) <= 1e-12 * np.linalg.norm(want)
    assert len(seen_parts) > 1


def test_quadratic_groups_are_planned_and_evaluated():
    """Ising/Heisenberg + Z X Z decorations: the heavy groups (many <= 2-local z-patterns) take the
    separable-table path (variant 62); the interpreter rebuilds the tables and checks them against
    the direct pattern sum."""
    from qiskit_aqua_b200 import workloads as W
    for rb, L in ((3, 11), (4, 11), (3, 12), (4, 13)):
        n = 14
        _, x, z, ypow, c = W.ising_heisenberg(n, n_triples=900, seed=5)
        eng = PauliSumEngine(n, x, z, ypow, c, options={'tile_bits': L, 'reg_bits': rb})
        passes = pi.export_part(eng, 0)
        nq = sum(int(np.count_nonzero(((ps['groups']['meta'] >> 22) & 63) == pi.VAR_QUAD)) for ps in passes)
        assert nq >= n                                  # every single-X group + the diagonal one
        assert sum(int(ch['nq']) for ps in passes for ch in ps['chunks']) == nq
        rng = np.random.default_rng(rb * 100 + L)
        psi = random_state(n, rng)
        y = np.zeros(1 << n, dtype=np.complex128)
        e = pi.run_part(passes, psi, y)
        want = po.mf_apply(n, x, z, ypow, np.asarray(c, dtype=np.complex128), psi)
        assert np.linalg.norm(y - want) <= 1e-12 * np.linalg.norm(want)
        eh = pi.run_part(passes, psi, None, herm=True)
        assert abs(eh.real - np.vdot(psi, want).real) <= 1e-11 * max(1.0, abs(np.vdot(psi, want)))


@pytest.mark.parametrize('rb', [3, 4])
@pytest.mark.parametrize('kind', ['c4_style', 'molecular', 'c5_style'])
def test_bench_workload_shapes(kind, rb):
    """The benchmark operator families at 13-14 qubits: TFIM + Heisenberg + 3-local part (C4),
    Jordan-Wigner molecular-style sum (C3), random weight <= 4 Paulis (C5)."""
    from qiskit_aqua_b200 import workloads as W
    if kind == 'c4_style':
        n = 13
        _, x, z, ypow, c = W.ising_heisenberg(n, n_triples=1500, seed=30)
    elif kind == 'molecular':
        _, x, z, ypow, c = W.molecular_style(12)
        n = 12
    else:
        n = 14
        _, x, z, ypow, c = W.random_weighted_paulis(n, 2000, 4, seed=34)
    eng = PauliSumEngine(n, x, z, ypow, c, options={'tile_bits': 11, 'reg_bits': rb})
    info = eng.info()
    assert info['n_bundles'] < info['n_group_records']            # partner sets are shared
    passes = pi.export_part(eng, 0)
    rng = np.random.default_rng(3)
    psi = random_state(n, rng)
    y = np.zeros(1 << n, dtype=np.complex128)
    pi.run_part(passes, psi, y)
    want = po.mf_apply(n, x, z, ypow, np.asarray(c, dtype=np.complex128), psi)
    assert np.linalg.norm(y - want) <= 1e-12 * np.linalg.norm(want)
    if eng.info()['is_hermitian']:
        e = pi.run_part(passes, psi, None, herm=True)
        assert abs(e.real - np.vdot(psi, want).real) <= 1e-11 * max(1.0, abs(np.vdot(psi, want)))


@pytest.mark.parametrize('n,T,late,split', [(14, 500, 2, 1), (15, 700, 3, 1), (14, 500, 2, 0)])
def test_host_pipeline_schedule_only_reads_arrived_chunks(n, T, late, split):
    """psum_expect_host / run_ready (psum.cu): the state arrives in 2^late chunks in index order; after
    chunk k every pass launches the tile range whose chunks are all present.  Emulated here with the
    exported plan: amplitudes that have not arrived are NaN, so a range launched too early poisons the
    sum; every tile must be launched exactly once and the total must equal <psi|H|psi>."""
    rng = np.random.default_rng(31 * n + T + late)
    x, z, ypow, c = random_terms(n, T, rng, complex_weights=False)
    eng = PauliSumEngine(n, x, z, ypow, c, options={'tile_bits': 11, 'late_bits': late, 'late_split': split})
    passes = pi.export_part(eng, 0)
    cbits = n - late
    lmask = (1 << late) - 1
    psi = random_state(n, rng)
    arrived = np.full(1 << n, np.nan + 0j)
    total = 0.0 + 0.0j
    launched = [0] * len(passes)
    n_progressive = 0
    for k in range(1 << late):
        arrived[k << cbits:(k + 1) << cbits] = psi[k << cbits:(k + 1) << cbits]
        for p, ps in enumerate(passes):
            free = 0
            for j in range(ps['L']):
                free |= 1 << int(ps['fbit'][j])
            T_ = lmask if eng.pass_info(p)['n_remote_groups'] else (free >> cbits) & lmask
            if (k & T_) != T_:
                continue
            u, j = 0, 0
            for b in range(late):
                if not (T_ >> b) & 1:
                    u |= ((k >> b) & 1) << j
                    j += 1
            per = (1 << ps['n_nonfree']) >> j
            total += pi.run_pass(ps, arrived, None, tiles=(u * per, (u + 1) * per))
            launched[p] += per
            if T_ and k != lmask:
                n_progressive += 1
    assert launched == [1 << ps['n_nonfree'] for ps in passes]
    want = np.vdot(psi, po.mf_apply(n, x, z, ypow, c, psi))
    assert np.isfinite(total.real) and abs(total - want) <= 1e-11 * max(1.0, abs(want))
    if split:
        assert n_progressive > 0        # some late ranges ran before the last chunk arrived
