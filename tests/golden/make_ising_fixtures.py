"""Runs the REFERENCE'S OWN Ising-problem generators (qiskit/optimization/applications/ising/*.py)
and its NumPyMinimumEigensolver on seeded instances -- including the instances of the reference's
own tests (test/optimization/test_{clique,exact_cover,graph_partition,knapsack,partition,
set_packing,stable_set,tsp,vehicle_routing,vertex_cover}.py) -- and records operators, shifts,
helper values and ground states in tests/golden/reference_ising.json.

    python tests/golden/make_ising_fixtures.py        (needs /root/reference; CPU only)

Same arrangement as make_reference_fixtures.py: terra is stubbed (tests/golden/terra_stub.py), the
reference's package __init__ files are bypassed, the listed reference modules run unmodified.
"""
import importlib
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_fixtures as mrf  # noqa: E402

SUBSETS_EXACT = [[2, 3, 4], [1, 2], [3, 4], [1, 2, 3]]     # test/optimization/sample.exactcover
SUBSETS_PACK = [[4, 5], [4], [5]]                          # test/optimization/sample.setpacking
NUMBERS = [8, 12, 4]                                       # test/optimization/sample.partition


def cplx(v):
    v = complex(v)
    return [v.real, v.imag]


def paulis_of(op):
    return [[cplx(w), p.to_label()] for w, p in op.paulis]


def bitstrings(n):
    return [np.array(b) for b in itertools.product((0, 1), repeat=n)]


def main():
    R = mrf.load_reference()
    ag = R['aqua_globals']
    algs = sys.modules['qiskit.aqua.algorithms']
    algs.MinimumEigensolverResult = importlib.import_module(
        'qiskit.aqua.algorithms.minimum_eigen_solvers.minimum_eigen_solver').MinimumEigensolverResult
    mods = {name: importlib.import_module('qiskit.optimization.applications.ising.' + name)
            for name in ('clique', 'exact_cover', 'graph_partition', 'knapsack', 'partition', 'set_packing',
                         'stable_set', 'tsp', 'vehicle_routing', 'vertex_cover', 'max_cut')}
    common = R['common']
    solver = R['NumPyMinimumEigensolver']

    def ground(op):
        res = solver(op, aux_operators=[]).run()
        x = common.sample_most_likely(res.eigenstate.primitive.data
                                      if hasattr(res.eigenstate.primitive, 'data') else res.eigenstate)
        return {'eigenvalue': cplx(res.eigenvalue), 'x': [int(v) for v in x]}

    def graph(n, p, seed, weight_range=10):
        ag.random_seed = seed          # the reference tests seed the global and call without `seed`
        return common.random_graph(n, edge_prob=p, weight_range=weight_range)

    out = {'note': 'outputs of the reference Ising generators (terra Pauli stubbed); see make_ising_fixtures.py'}

    # ---- graph problems -------------------------------------------------------------------------
    for name, cases in (('graph_partition', [(4, 0.8, 100), (6, 0.5, 7)]),
                        ('vertex_cover', [(3, 0.8, 100), (6, 0.5, 11)]),
                        ('stable_set', [(5, 0.5, 8123179), (7, 0.4, 3)]),
                        ('max_cut', [(4, 0.5, 8123179999), (6, 0.5, 99)])):
        m = mods[name]
        out[name] = []
        for n, p, seed in cases:
            w = graph(n, p, seed)
            op, shift = m.get_operator(w)
            rec = {'n': n, 'edge_prob': p, 'seed': seed, 'w': w.tolist(), 'paulis': paulis_of(op),
                   'shift': float(shift), 'ground': ground(op)}
            xs = bitstrings(n)
            if name == 'graph_partition':
                rec['objective_value'] = [float(m.objective_value(x, w)) for x in xs]
            elif name == 'vertex_cover':
                rec['check_full_edge_coverage'] = [bool(m.check_full_edge_coverage(x, w)) for x in xs]
            elif name == 'stable_set':
                rec['stable_set_value'] = [[float(a), bool(b)] for a, b in (m.stable_set_value(x, w) for x in xs)]
            else:
                rec['max_cut_value'] = [float(m.max_cut_value(x, w)) for x in xs]
            out[name].append(rec)
    out['clique'] = []
    for n, p, seed, K in ((5, 0.8, 100, 5), (5, 0.8, 100, 3), (6, 0.6, 21, 3)):
        w = graph(n, p, seed)
        op, shift = mods['clique'].get_operator(w, K)
        out['clique'].append({'n': n, 'edge_prob': p, 'seed': seed, 'K': K, 'w': w.tolist(), 'paulis': paulis_of(op),
                              'shift': float(shift), 'ground': ground(op),
                              'satisfy_or_not': [bool(mods['clique'].satisfy_or_not(x, w, K)) for x in bitstrings(n)]})

    # ---- set problems ---------------------------------------------------------------------------
    out['exact_cover'] = []
    for subsets in (SUBSETS_EXACT, [[1, 2], [3], [2, 3], [1], [1, 3, 4], [4]]):
        op, shift = mods['exact_cover'].get_operator(subsets)
        out['exact_cover'].append({'subsets': subsets, 'paulis': paulis_of(op), 'shift': float(shift), 'ground': ground(op),
                                   'check_solution_satisfiability': [bool(mods['exact_cover'].check_solution_satisfiability(x, subsets))
                                                                     for x in bitstrings(len(subsets))]})
    out['set_packing'] = []
    for subsets in (SUBSETS_PACK, [[1, 2], [2, 3], [4], [3, 5], [5, 6, 1]]):
        op, shift = mods['set_packing'].get_operator(subsets)
        out['set_packing'].append({'subsets': subsets, 'paulis': paulis_of(op), 'shift': float(shift), 'ground': ground(op),
                                   'check_disjoint': [bool(mods['set_packing'].check_disjoint(x, subsets))
                                                      for x in bitstrings(len(subsets))]})
    out['partition'] = []
    ag.random_seed = 5
    for numbers in (np.array(NUMBERS), common.random_number_list(7, weight_range=25)):
        op, shift = mods['partition'].get_operator(numbers)
        out['partition'].append({'numbers': [int(v) for v in numbers], 'paulis': paulis_of(op), 'shift': float(shift),
                                 'ground': ground(op),
                                 'partition_value': [float(mods['partition'].partition_value(x, numbers))
                                                     for x in bitstrings(len(numbers))]})
    out['random_number_list'] = {'seed': 5, 'n': 7, 'weight_range': 25, 'numbers': out['partition'][1]['numbers']}

    # ---- TSP ------------------------------------------------------------------------------------
    out['tsp'] = []
    for n, seed, penalty in ((3, 80598, 1e5), (3, 12, 250.0), (4, 2, 1e5)):
        ag.random_seed = seed
        ins = mods['tsp'].random_tsp(n)
        op, shift = mods['tsp'].get_operator(ins, penalty)
        rec = {'n': n, 'seed': seed, 'penalty': penalty, 'coord': ins.coord.tolist(), 'w': ins.w.tolist(),
               'paulis': paulis_of(op), 'shift': float(shift)}
        if n == 3:
            g = ground(op)
            rec['ground'] = g
            rec['order'] = mods['tsp'].get_tsp_solution(np.array(g['x']))
            rec['tsp_value'] = float(mods['tsp'].tsp_value(rec['order'], ins.w))
            rec['tsp_feasible'] = [bool(mods['tsp'].tsp_feasible(x)) for x in bitstrings(9)]
        out['tsp'].append(rec)
    out['tsp_get_solution'] = [[[1, 0, 0, 0, 1, 0, 0, 0, 1], mods['tsp'].get_tsp_solution([1, 0, 0, 0, 1, 0, 0, 0, 1])],
                               [[1, 0, 0, 1, 1, 0, 0, 0, 0], mods['tsp'].get_tsp_solution([1, 0, 0, 1, 1, 0, 0, 0, 0])]]

    # ---- knapsack -------------------------------------------------------------------------------
    out['knapsack'] = []
    for values, weights, max_weight in (([10, 40, 50, 75], [5, 10, 3, 12], 20), ([10, 40, 50, 75], [5, 10, 3, 12], 0),
                                        ([10, 40, 50, 75], [5, 10, 3, 12], 1000), ([3, 1, 4], [2, 7, 1], 8)):
        op, shift = mods['knapsack'].get_operator(values, weights, max_weight)
        g = ground(op)
        sol = mods['knapsack'].get_solution(np.array(g['x']), values)
        val, wt = mods['knapsack'].knapsack_value_weight(sol, values, weights)
        out['knapsack'].append({'values': values, 'weights': weights, 'max_weight': max_weight, 'paulis': paulis_of(op),
                                'shift': float(shift), 'ground': g, 'solution': [int(v) for v in sol],
                                'value': float(val), 'weight': float(wt)})

    # ---- vehicle routing ------------------------------------------------------------------------
    out['vehicle_routing'] = []
    inst2 = np.zeros((2, 2))
    inst2[0, 1] = inst2[1, 0] = 0.8
    r = np.random.default_rng(4)
    inst3 = np.rint(r.uniform(1, 20, (3, 3)))
    inst3 = inst3 + inst3.T
    np.fill_diagonal(inst3, 0)
    for inst, n, K in ((inst2, 2, 1), (inst3, 3, 2)):
        vr = mods['vehicle_routing']
        Q, g_, c = vr.get_vehiclerouting_matrices(inst, n, K)
        op = vr.get_operator(inst, n, K)
        N = n * (n - 1)
        out['vehicle_routing'].append({'instance': inst.tolist(), 'n': n, 'K': K, 'Q': Q.tolist(), 'g': g_.tolist(), 'c': float(c),
                                       'paulis': paulis_of(op), 'ground': ground(op),
                                       'cost': [float(vr.get_vehiclerouting_cost(inst, n, K, x)) for x in bitstrings(N)]})

    # ---- file formats ---------------------------------------------------------------------------
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        ag.random_seed = 31
        path = os.path.join(d, 'g.gset')
        w = common.random_graph(5, edge_prob=0.6, weight_range=10, savefile=path)
        text = open(path).read()
        # the reference writes float weights ("1 2 3.0") which its own parser cannot read back
        # (int('3.0')); the parser fixture therefore uses an integer-weight file
        ints = '\n'.join(' '.join(str(int(float(t))) for t in line.split()) for line in text.strip().split('\n')) + '\n'
        open(path, 'w').write(ints)
        out['gset'] = {'seed': 31, 'n': 5, 'edge_prob': 0.6, 'w': w.tolist(), 'written': text, 'int_file': ints,
                       'parsed': common.parse_gset_format(path).tolist(),
                       'gset_result': {str(k): int(v) for k, v in common.get_gset_result(np.array([1, 0, 0, 1, 1])).items()}}
        ag.random_seed = 8
        path = os.path.join(d, 't.tsp')
        ins = mods['tsp'].random_tsp(4, savefile=path, name='fx')
        back = mods['tsp'].parse_tsplib_format(path)
        out['tsplib'] = {'seed': 8, 'n': 4, 'coord': ins.coord.tolist(), 'written': open(path).read(),
                         'parsed_w': back.w.tolist(), 'parsed_coord': back.coord.tolist()}
        path = os.path.join(d, 'n.txt')
        ag.random_seed = 9
        nums = common.random_number_list(5, weight_range=50, savefile=path)
        out['numbers_file'] = {'seed': 9, 'numbers': [int(v) for v in nums], 'written': open(path).read(),
                               'read': [int(v) for v in common.read_numbers_from_file(path)]}

    path = os.path.join(HERE, 'reference_ising.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=0)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
