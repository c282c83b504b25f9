"""Ising-problem generators (qiskit_aqua_b200/ising.py) against outputs of the reference's own
modules (tests/golden/reference_ising.json, produced by tests/golden/make_ising_fixtures.py from
qiskit/optimization/applications/ising/*.py), and -- on the GPU -- the reference's own test
assertions (test/optimization/test_*.py) through NumPyMinimumEigensolver's device diagonal branch.
"""
import itertools
import json
import os

import numpy as np
import pytest

import qiskit_aqua_b200 as q
from qiskit_aqua_b200 import ising

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, 'golden', 'reference_ising.json')))


def bitstrings(n):
    return [np.array(b) for b in itertools.product((0, 1), repeat=n)]


def assert_same_operator(op, rec_paulis, tol=1e-12):
    """Same simplified `paulis` list as the reference: labels in the same order, weights equal."""
    got = [(p.to_label(), complex(w)) for w, p in op.paulis]
    want = [(lab, complex(*c)) for c, lab in rec_paulis]
    assert [g[0] for g in got] == [w[0] for w in want]
    scale = max(1.0, max(abs(w[1]) for w in want))
    for (_, a), (_, b) in zip(got, want):
        assert abs(a - b) <= tol * scale


def energy(op, x):
    """<x|H|x> of a diagonal operator for the bit string x (qubit 0 first)."""
    idx = int(sum(int(b) << k for k, b in enumerate(x)))
    return sum(complex(w).real * (-1) ** bin(p.z_mask & idx).count('1') for w, p in op.paulis)


GRAPH = {'graph_partition': ising.graph_partition, 'vertex_cover': ising.vertex_cover,
         'stable_set': ising.stable_set, 'max_cut': ising.max_cut}


@pytest.mark.parametrize('name', sorted(GRAPH))
def test_graph_problems_match_reference(name):
    mod = GRAPH[name]
    for rec in FX[name]:
        q.aqua_globals.random_seed = rec['seed']
        w = ising.random_graph(rec['n'], edge_prob=rec['edge_prob'], weight_range=10)
        assert np.array_equal(w, np.array(rec['w']))
        op, shift = mod.get_operator(w)
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], abs=1e-12)
        xs = bitstrings(rec['n'])
        if name == 'graph_partition':
            assert [float(mod.objective_value(x, w)) for x in xs] == rec['objective_value']
        elif name == 'vertex_cover':
            assert [mod.check_full_edge_coverage(x, w) for x in xs] == rec['check_full_edge_coverage']
        elif name == 'stable_set':
            assert [[float(a), bool(b)] for a, b in (mod.stable_set_value(x, w) for x in xs)] == rec['stable_set_value']
        else:
            assert [float(mod.max_cut_value(x, w)) for x in xs] == rec['max_cut_value']
        # the recorded reference ground state is a minimum of OUR operator too
        e = [energy(op, x[::-1]) for x in xs]          # bitstrings() lists qubit n-1 ... 0
        assert min(e) == pytest.approx(rec['ground']['eigenvalue'][0], abs=1e-9)
        assert energy(op, rec['ground']['x']) == pytest.approx(min(e), abs=1e-9)


def test_clique_matches_reference():
    for rec in FX['clique']:
        w = np.array(rec['w'])
        op, shift = ising.clique.get_operator(w, rec['K'])
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], rel=1e-13)
        assert [bool(ising.clique.satisfy_or_not(x, w, rec['K'])) for x in bitstrings(rec['n'])] == rec['satisfy_or_not']


def test_set_problems_match_reference():
    for rec in FX['exact_cover']:
        op, shift = ising.exact_cover.get_operator(rec['subsets'])
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], abs=1e-12)
        got = [ising.exact_cover.check_solution_satisfiability(x, rec['subsets']) for x in bitstrings(len(rec['subsets']))]
        assert got == rec['check_solution_satisfiability']
    for rec in FX['set_packing']:
        op, shift = ising.set_packing.get_operator(rec['subsets'])
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], abs=1e-12)
        got = [ising.set_packing.check_disjoint(x, rec['subsets']) for x in bitstrings(len(rec['subsets']))]
        assert got == rec['check_disjoint']


def test_partition_and_number_list_match_reference():
    rl = FX['random_number_list']
    q.aqua_globals.random_seed = rl['seed']
    assert [int(v) for v in ising.random_number_list(rl['n'], weight_range=rl['weight_range'])] == rl['numbers']
    for rec in FX['partition']:
        numbers = np.array(rec['numbers'])
        op, shift = ising.partition.get_operator(numbers)
        assert_same_operator(op, rec['paulis'])
        assert shift == rec['shift']
        assert [float(ising.partition.partition_value(x, numbers)) for x in bitstrings(len(numbers))] == rec['partition_value']


def test_tsp_matches_reference():
    for rec in FX['tsp']:
        q.aqua_globals.random_seed = rec['seed']
        ins = ising.tsp.random_tsp(rec['n'])
        assert np.allclose(ins.coord, np.array(rec['coord']), rtol=0, atol=0)
        assert np.array_equal(ins.w, np.array(rec['w']))
        op, shift = ising.tsp.get_operator(ins, rec['penalty'])
        assert op.num_qubits == rec['n'] ** 2
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], rel=1e-13)
        if 'ground' in rec:
            assert ising.tsp.get_tsp_solution(np.array(rec['ground']['x'])) == rec['order']
            assert ising.tsp.tsp_value(rec['order'], ins.w) == rec['tsp_value']
            assert [ising.tsp.tsp_feasible(x) for x in bitstrings(9)] == rec['tsp_feasible']
    for x, want in FX['tsp_get_solution']:      # test/optimization/test_tsp.py:47-52
        assert ising.tsp.get_tsp_solution(x) == want


def test_knapsack_matches_reference():
    for rec in FX['knapsack']:
        op, shift = ising.knapsack.get_operator(rec['values'], rec['weights'], rec['max_weight'])
        assert_same_operator(op, rec['paulis'])
        assert shift == pytest.approx(rec['shift'], rel=1e-13)
        sol = ising.knapsack.get_solution(np.array(rec['ground']['x']), rec['values'])
        assert [int(v) for v in sol] == rec['solution']
        value, weight = ising.knapsack.knapsack_value_weight(sol, rec['values'], rec['weights'])
        assert (value, weight) == (rec['value'], rec['weight'])
    for bad in (([1, 2], [1], 3), ([1, -2], [1, 1], 3), ([0, 0], [1, 1], 3), ([1, 2], [1, 1], -1)):
        with pytest.raises(ValueError):       # knapsack.py:75-82
            ising.knapsack.get_operator(*bad)


def test_vehicle_routing_matches_reference():
    for rec in FX['vehicle_routing']:
        inst, n, K = np.array(rec['instance']), rec['n'], rec['K']
        Q, g, c = ising.vehicle_routing.get_vehiclerouting_matrices(inst, n, K)
        assert np.allclose(Q, np.array(rec['Q']), rtol=1e-13, atol=0)
        assert np.allclose(g, np.array(rec['g']), rtol=1e-13, atol=1e-12)
        assert c == pytest.approx(rec['c'], rel=1e-13)
        op = ising.vehicle_routing.get_operator(inst, n, K)
        assert_same_operator(op, rec['paulis'], tol=1e-12)
        cost = [float(ising.vehicle_routing.get_vehiclerouting_cost(inst, n, K, x)) for x in bitstrings(n * (n - 1))]
        assert np.allclose(cost, rec['cost'], rtol=1e-12, atol=1e-9)
    # test/optimization/test_vehicle_routing.py:40-54: Z0, Z1, identity with weights 79.6, 79.6, 160.8
    op = ising.vehicle_routing.get_operator(np.array([[0, 0.8], [0.8, 0]]), 2, 1)
    assert [p.to_label() for _, p in op.paulis] == ['IZ', 'ZI', 'II']
    assert np.allclose([complex(w).real for w, _ in op.paulis], [79.6, 79.6, 160.8], rtol=1e-2)


def test_file_formats_match_reference(tmp_path):
    g = FX['gset']
    q.aqua_globals.random_seed = g['seed']
    path = str(tmp_path / 'g.gset')
    w = ising.random_graph(g['n'], edge_prob=g['edge_prob'], weight_range=10, savefile=path)
    assert np.array_equal(w, np.array(g['w']))
    assert open(path).read() == g['written']
    open(path, 'w').write(g['int_file'])
    assert np.array_equal(ising.parse_gset_format(path), np.array(g['parsed']))
    assert {str(k): int(v) for k, v in ising.get_gset_result(np.array([1, 0, 0, 1, 1])).items()} == g['gset_result']
    t = FX['tsplib']
    q.aqua_globals.random_seed = t['seed']
    path = str(tmp_path / 't.tsp')
    ins = ising.tsp.random_tsp(t['n'], savefile=path, name='fx')
    assert open(path).read() == t['written']
    back = ising.tsp.parse_tsplib_format(path)
    assert np.array_equal(back.w, np.array(t['parsed_w']))
    assert np.array_equal(back.coord, np.array(t['parsed_coord']))
    assert np.array_equal(ins.coord, np.array(t['coord']))
    nf = FX['numbers_file']
    q.aqua_globals.random_seed = nf['seed']
    path = str(tmp_path / 'n.txt')
    assert [int(v) for v in ising.random_number_list(5, weight_range=50, savefile=path)] == nf['numbers']
    assert open(path).read() == nf['written']
    assert [int(v) for v in ising.read_numbers_from_file(path)] == nf['read']


def test_sample_most_likely_forms():
    assert list(ising.sample_most_likely({'011': 5, '100': 9})) == [0, 0, 1]      # counts: qubit 0 is the last char
    v = np.zeros(8)
    v[6] = -0.9
    v[1] = 0.3
    assert list(ising.sample_most_likely(v)) == [0, 1, 1]


def test_aqua_globals_seed_semantics():
    q.aqua_globals.random_seed = 17
    a = q.aqua_globals.random.random()
    q.aqua_globals.random_seed = 17
    assert q.aqua_globals.random.random() == a
    assert q.aqua_globals.random.random() != a
    w1 = ising.random_graph(5, seed=3)
    w2 = ising.random_graph(5, seed=3)
    assert np.array_equal(w1, w2)


# ---------------------------------------------------------------------------------------------
# GPU: the reference's own test assertions through the device diagonal branch
# ---------------------------------------------------------------------------------------------
def _solve(op):
    res = q.NumPyMinimumEigensolver(op, aux_operators=[]).run()
    return res, ising.sample_most_likely(res.eigenstate)


def _all_ground_records():
    for name in ('graph_partition', 'vertex_cover', 'stable_set', 'max_cut', 'clique', 'exact_cover',
                 'set_packing', 'partition', 'knapsack', 'vehicle_routing'):
        for k, rec in enumerate(FX[name]):
            yield pytest.param(name, k, id='%s-%d' % (name, k))
    for k, rec in enumerate(FX['tsp']):
        if 'ground' in rec:
            yield pytest.param('tsp', k, id='tsp-%d' % k)


def _operator_of(name, rec):
    lab = [(complex(*c), l) for c, l in rec['paulis']]
    return q.WeightedPauliOperator(paulis=[[w, q.Pauli(l)] for w, l in lab])


@pytest.mark.gpu
@pytest.mark.parametrize('name,k', list(_all_ground_records()))
def test_device_ground_state_matches_reference_solver(name, k):
    """Eigenvalue equal to the reference's NumPyMinimumEigensolver; the returned basis state is a
    minimiser (equal to the reference's when the minimum is unique)."""
    rec = FX[name][k]
    op = _operator_of(name, rec)
    res, x = _solve(op)
    want = rec['ground']['eigenvalue'][0]
    assert res.eigenvalue.real == pytest.approx(want, rel=1e-12, abs=1e-9)
    assert energy(op, x) == pytest.approx(want, rel=1e-12, abs=1e-9)
    n = op.num_qubits
    if n <= 12:
        e = np.array([energy(op, [(i >> b) & 1 for b in range(n)]) for i in range(1 << n)])
        if np.sum(np.isclose(e, e.min(), rtol=1e-12, atol=1e-9)) == 1:
            assert [int(v) for v in x] == rec['ground']['x']


@pytest.mark.gpu
def test_reference_test_assertions_on_device():
    # test_clique.py:58-66
    q.aqua_globals.random_seed = 100
    w = ising.random_graph(5, edge_prob=0.8, weight_range=10)
    op, _ = ising.clique.get_operator(w, 5)
    _, x = _solve(op)
    sol = ising.clique.get_graph_solution(x)
    np.testing.assert_array_equal(sol, [1, 1, 1, 1, 1])
    brute = any(ising.clique.satisfy_or_not(b, w, 5) for b in bitstrings(5))
    assert ising.clique.satisfy_or_not(sol, w, 5) == brute
    # test_exact_cover.py:54-64
    subsets = [[2, 3, 4], [1, 2], [3, 4], [1, 2, 3]]
    op, _ = ising.exact_cover.get_operator(subsets)
    _, x = _solve(op)
    sol = ising.exact_cover.get_solution(x)
    np.testing.assert_array_equal(sol, [0, 1, 1, 0])
    assert ising.exact_cover.check_solution_satisfiability(sol, subsets)
    # test_set_packing.py:62-71
    subsets = [[4, 5], [4], [5]]
    op, _ = ising.set_packing.get_operator(subsets)
    _, x = _solve(op)
    sol = ising.set_packing.get_solution(x)
    np.testing.assert_array_equal(sol, [0, 1, 1])
    best = max(int(np.count_nonzero(b)) for b in bitstrings(3) if ising.set_packing.check_disjoint(b, subsets))
    assert np.count_nonzero(sol) == best
    # test_partition.py:39-47
    op, _ = ising.partition.get_operator(np.array([8, 12, 4]))
    _, x = _solve(op)
    if x[0] != 0:
        x = np.logical_not(x) * 1
    np.testing.assert_array_equal(x, [0, 1, 0])
    # test_stable_set.py:40-50
    q.aqua_globals.random_seed = 8123179
    w = ising.random_graph(5, edge_prob=0.5)
    op, offset = ising.stable_set.get_operator(w)
    res, x = _solve(op)
    assert res.eigenvalue.real == pytest.approx(-2.5)
    assert res.eigenvalue.real + offset == pytest.approx(-3)
    np.testing.assert_array_equal(ising.stable_set.get_graph_solution(x), [1, 0, 1, 0, 1])
    assert ising.stable_set.stable_set_value(x, w) == (3, True)
    # test_vertex_cover.py:60-69
    q.aqua_globals.random_seed = 100
    w = ising.random_graph(3, edge_prob=0.8, weight_range=10)
    op, _ = ising.vertex_cover.get_operator(w)
    _, x = _solve(op)
    sol = ising.vertex_cover.get_graph_solution(x)
    np.testing.assert_array_equal(sol, [0, 0, 1])
    best = min(int(np.count_nonzero(b)) for b in bitstrings(3) if ising.vertex_cover.check_full_edge_coverage(b, w))
    assert np.count_nonzero(sol) == best
    # test_graph_partition.py:57-70
    q.aqua_globals.random_seed = 100
    w = ising.random_graph(4, edge_prob=0.8, weight_range=10)
    op, _ = ising.graph_partition.get_operator(w)
    _, x = _solve(op)
    sol = ising.graph_partition.get_graph_solution(x)
    assert ising.graph_partition.objective_value(np.array([0, 1, 0, 1]), w) == ising.graph_partition.objective_value(sol, w)
    best = min(ising.graph_partition.objective_value(b, w) for b in bitstrings(4) if 2 * np.count_nonzero(b) == 4)
    assert ising.graph_partition.objective_value(x, w) == best
    # test_tsp.py:36-45 (9 qubits)
    q.aqua_globals.random_seed = 80598
    ins = ising.tsp.random_tsp(3)
    op, _ = ising.tsp.get_operator(ins)
    res = q.NumPyMinimumEigensolver(op).run()
    order = ising.tsp.get_tsp_solution(ising.sample_most_likely(res.eigenstate))
    assert ising.tsp.tsp_value(order, ins.w) == ising.tsp.tsp_value([1, 2, 0], ins.w)
    # test_knapsack.py:41-72 (9, 5 and 14 qubits)
    for max_weight, want, value, weight in ((20, [1, 0, 1, 1], 135, 20), (0, [0, 0, 0, 0], 0, 0), (1000, [1, 1, 1, 1], 175, 30)):
        values, weights = [10, 40, 50, 75], [5, 10, 3, 12]
        op, _ = ising.knapsack.get_operator(values, weights, max_weight)
        res = q.NumPyMinimumEigensolver(op).run()
        sol = ising.knapsack.get_solution(ising.sample_most_likely(res.eigenstate), values)
        np.testing.assert_array_equal(sol, want)
        assert ising.knapsack.knapsack_value_weight(sol, values, weights) == (value, weight)
    # test_vehicle_routing.py:56-61
    op = ising.vehicle_routing.get_operator(np.array([[0, 0.8], [0.8, 0]]), 2, 1)
    res = q.NumPyMinimumEigensolver(op).run()
    np.testing.assert_array_almost_equal(np.array([0., 0., 0., 1.]), np.abs(res.eigenstate.to_matrix()) ** 2, 4)
    np.testing.assert_array_equal(ising.vehicle_routing.get_vehiclerouting_solution(None, 2, 1, res), [1, 1])
