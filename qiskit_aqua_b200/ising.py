"""Ising-problem generators feeding the hot path: each returns the diagonal Pauli-sum Hamiltonian
that `NumPyMinimumEigensolver` then solves with the device diagonal kernel (`psum_diag_kmin`).

Drop-in for `qiskit/optimization/applications/ising/*.py`.  The reference
builds every Pauli with its own pair of length-n bool arrays inside nested Python loops; here a
generator describes the Hamiltonian as a few *blocks* of Z-masks + weights produced with numpy
index arithmetic, emitted in the reference's order of first appearance so that the simplified
operator (`WeightedPauliOperator.__init__` merges duplicates and drops zeros,
legacy/weighted_pauli_operator.py:66-70) has the same `paulis` list, term by term.

Namespaces mirror the reference's module names:
    max_cut, graph_partition, clique, vertex_cover, stable_set, exact_cover, set_packing,
    partition, tsp, knapsack, vehicle_routing   (+ the helpers of ising/common.py)
"""
import math
from collections import OrderedDict, namedtuple

import numpy as np

from .aqua_globals import aqua_globals
from .operators import Pauli, WeightedPauliOperator


# ---------------------------------------------------------------------------------------------
# term blocks
# ---------------------------------------------------------------------------------------------
def _bit(q):
    return np.left_shift(np.uint64(1), np.asarray(q, dtype=np.uint64))


class _ZTerms:
    """Ordered sequence of (z_mask, weight) terms of a diagonal Hamiltonian."""

    def __init__(self, num_qubits):
        if num_qubits > 64:
            raise ValueError('at most 64 qubits')
        self.n = int(num_qubits)
        self._masks = []
        self._weights = []

    def add(self, masks, weights):
        """One block; `masks`/`weights` broadcast against each other, C order = emission order.
        A 2-D block [m, k] therefore emits the k terms of row 0, then those of row 1, ..."""
        masks = np.asarray(masks, dtype=np.uint64)
        weights = np.asarray(weights, dtype=np.float64)
        masks, weights = np.broadcast_arrays(masks, weights)
        self._masks.append(masks.reshape(-1))
        self._weights.append(weights.reshape(-1))

    def operator(self):
        masks = np.concatenate(self._masks) if self._masks else np.zeros(0, np.uint64)
        weights = np.concatenate(self._weights) if self._weights else np.zeros(0)
        n = self.n
        return WeightedPauliOperator(paulis=[[w, Pauli.from_masks(0, m, n)]
                                             for w, m in zip(weights.tolist(), masks.tolist())])


def _lower_pairs(n, keep=None):
    """(i, j) with j < i in the order `for i: for j in range(i)`; optionally filtered."""
    i, j = np.tril_indices(n, -1)
    if keep is not None:
        sel = keep[i, j]
        i, j = i[sel], j[sel]
    return i, j


def _pair_with_singles(i, j, w_pair, w_i, w_j):
    """Block rows [Z_i Z_j, Z_i, Z_j] per pair (the reference's edge-loop body)."""
    masks = np.stack([_bit(i) | _bit(j), _bit(i), _bit(j)], axis=1)
    ones = np.ones(len(i))
    weights = np.stack([w_pair * ones, w_i * ones, w_j * ones], axis=1)
    return masks, weights


# ---------------------------------------------------------------------------------------------
# ising/common.py
# ---------------------------------------------------------------------------------------------
def random_graph(n, weight_range=10, edge_prob=0.3, negative_weight=True, savefile=None, seed=None):
    """Erdos-Renyi graph with the reference's draw order (random <= p, integers, random >= .5);
    draws from `aqua_globals.random`, re-seeded when `seed` is given (common.py:23-60)."""
    assert weight_range >= 0
    if seed:
        aqua_globals.random_seed = seed
    rng = aqua_globals.random
    w = np.zeros((n, n))
    # the three draws of an edge are conditional on each other, so the pairs are visited one by one in
    # the reference's order (row-major upper triangle); everything after the draws is array code
    for i, j in zip(*np.triu_indices(n, k=1)):
        if rng.random() > edge_prob:
            continue
        weight = rng.integers(1, weight_range)
        w[i, j] = -weight if (rng.random() >= 0.5 and negative_weight) else weight
    rows, cols = np.nonzero(w)                      # the edges (weights are >= 1): upper triangle, row-major
    w += w.T
    if savefile:
        lines = ['%d %d\n' % (n, len(rows))]
        lines += ['%d %d %s\n' % (a + 1, b + 1, w[a, b]) for a, b in zip(rows, cols)]
        with open(savefile, 'w', encoding='utf8') as handle:
            handle.writelines(lines)
    return w


def random_number_list(n, weight_range=100, savefile=None, seed=None):
    """n integers in [1, weight_range] from `aqua_globals.random` (common.py:63-83)."""
    if seed:
        aqua_globals.random_seed = seed
    numbers = aqua_globals.random.integers(low=1, high=(weight_range + 1), size=n)
    if savefile:
        with open(savefile, 'w', encoding='utf8') as f:
            f.writelines('{}\n'.format(v) for v in numbers)
    return numbers


def read_numbers_from_file(filename):
    """One integer per line (common.py:86-101)."""
    numbers = []
    with open(filename, encoding='utf8') as f:
        for line in f:
            value = float(line)
            assert int(round(value)) == value
            numbers.append(int(round(value)))
    return np.array(numbers)


def parse_gset_format(filename):
    """Gset file -> symmetric matrix (common.py:104-131).  As in the reference the stored entry is
    the 0-based END-NODE INDEX of the edge, not the weight column of the file (reference quirk,
    `w[s, t] = t`; kept so that both read the same matrix)."""
    with open(filename, encoding='utf8') as f:
        rows = [[int(e) for e in line.split()] for line in f]
    n, m = rows[0]
    w = np.zeros((n, n))
    edges = np.array(rows[1:], dtype=np.int64).reshape(-1, 3)
    assert m == len(edges)
    for s, t, _ in edges:          # later duplicates overwrite earlier ones, as in the reference
        w[s - 1, t - 1] = t - 1
    w += w.T
    return w


def get_gset_result(x):
    """Binary string -> {1-based node: side} (common.py:134-143)."""
    return {i + 1: 1 - x[i] for i in range(len(x))}


def sample_most_likely(state_vector):
    """Most likely bit string of counts, a state function, a numpy vector or a device tensor
    (common.py:146-170); qubit 0 first."""
    if isinstance(state_vector, (OrderedDict, dict)):
        binary_string = sorted(state_vector.items(), key=lambda kv: kv[1])[-1][0]
        return np.asarray([int(y) for y in reversed(list(binary_string))])
    if hasattr(state_vector, 'device_vector'):
        state_vector = state_vector.device_vector()
    if hasattr(state_vector, 'is_cuda'):
        k = int(state_vector.abs().argmax().item())
        n = int(state_vector.numel()).bit_length() - 1
    else:
        n = int(np.log2(state_vector.shape[0]))
        k = int(np.argmax(np.abs(state_vector)))
    return ((k >> np.arange(n)) & 1).astype(float)


# ---------------------------------------------------------------------------------------------
# problems
# ---------------------------------------------------------------------------------------------
class max_cut:
    """ising/max_cut.py"""

    @staticmethod
    def get_operator(weight_matrix):
        """0.5 w_ij Z_i Z_j per edge, shift = -sum 0.5 w_ij (max_cut.py:31-54)."""
        w = np.asarray(weight_matrix)
        n = len(w)
        i, j = _lower_pairs(n, w != 0)
        terms = _ZTerms(n)
        terms.add(_bit(i) | _bit(j), 0.5 * w[i, j])
        shift = 0
        for v in w[i, j]:
            shift -= 0.5 * v
        return terms.operator(), shift

    @staticmethod
    def max_cut_value(x, w):
        return np.sum(w * np.outer(x, (1 - x)))

    @staticmethod
    def get_graph_solution(x):
        return 1 - x


class graph_partition:
    """ising/graph_partition.py: H = sum_edges (1 - Z_i Z_j)/2 + (sum_i Z_i)^2."""

    @staticmethod
    def get_operator(weight_matrix):
        w = np.asarray(weight_matrix)
        n = len(w)
        terms = _ZTerms(n)
        i, j = _lower_pairs(n, w != 0)
        terms.add(_bit(i) | _bit(j), -0.5)
        a, b = np.indices((n, n)).reshape(2, -1)      # every ordered pair, i != j (graph_partition.py:60-69)
        off = a != b
        terms.add(_bit(a[off]) | _bit(b[off]), 1.0)
        return terms.operator(), 0.5 * len(i) + n

    @staticmethod
    def objective_value(x, w):
        return np.sum(np.where(w != 0, 1, 0) * np.outer(x, (1 - x)))

    @staticmethod
    def get_graph_solution(x):
        return 1 - x


class clique:
    """ising/clique.py: H = A (K - sum x_v)^2 + K(K-1)/2 - sum_edges x_u x_v, x_v = (Z_v + 1)/2, A = 1000."""

    @staticmethod
    def get_operator(weight_matrix, K):  # noqa: N803
        w = np.asarray(weight_matrix)
        n = len(w)
        A = 1000
        Y = K - 0.5 * n
        terms = _ZTerms(n)
        a, b = np.indices((n, n)).reshape(2, -1)
        off = a != b
        terms.add(_bit(a[off]) | _bit(b[off]), A * 0.25)
        terms.add(_bit(np.arange(n)), -A * Y)
        i, j = _lower_pairs(n, w != 0)
        terms.add(*_pair_with_singles(i, j, -0.25, -0.25, -0.25))
        shift = A * Y * Y + n * A * 0.25 + 0.5 * K * (K - 1) - 0.25 * len(i)
        return terms.operator(), shift

    @staticmethod
    def satisfy_or_not(x, w, K):  # noqa: N803
        return np.sum(np.where(w != 0, 1, 0) * np.outer(x, x)) == K * (K - 1)

    @staticmethod
    def get_graph_solution(x):
        return 1 - x


class vertex_cover:
    """ising/vertex_cover.py: H = A sum_edges (1 - x_i)(1 - x_j) + sum_i Z_i terms, A = 5."""

    @staticmethod
    def get_operator(weight_matrix):
        w = np.asarray(weight_matrix)
        n = len(w)
        A = 5
        terms = _ZTerms(n)
        i, j = _lower_pairs(n, w != 0)
        terms.add(*_pair_with_singles(i, j, A * 0.25, -A * 0.25, -A * 0.25))
        terms.add(_bit(np.arange(n)), 0.5)
        return terms.operator(), A * 0.25 * len(i) + 0.5 * n

    @staticmethod
    def check_full_edge_coverage(x, w):
        x = np.asarray(x)
        uncovered = (np.asarray(w) != 0) & np.outer(x != 1, x != 1)
        return not bool(uncovered.any())

    @staticmethod
    def get_graph_solution(x):
        return 1 - x


class stable_set:
    """ising/stable_set.py (edge weights are ignored, only the pattern counts)."""

    @staticmethod
    def get_operator(w):
        w = np.asarray(w)
        n = len(w)
        terms = _ZTerms(n)
        i, j = np.triu_indices(n, 1)
        sel = w[i, j] != 0
        i, j = i[sel], j[sel]
        terms.add(_bit(i) | _bit(j), 0.5)
        degree = np.count_nonzero(w != 0, axis=1)
        terms.add(_bit(np.arange(n)), 0.5 - degree / 2)
        return terms.operator(), 0.5 * len(i) - n / 2

    @staticmethod
    def stable_set_value(x, w):
        x = np.asarray(x)
        assert len(x) == w.shape[0]
        clash = np.triu(np.asarray(w) != 0, 1) & np.outer(x == 1, x == 1)
        return np.sum(x), not bool(clash.any())

    @staticmethod
    def get_graph_solution(x):
        return x


def _membership(list_of_subsets):
    """universe (sorted unique elements), bool matrix [element, subset]."""
    universe = np.unique([e for sub in list_of_subsets for e in sub])
    member = np.array([[e in sub for sub in list_of_subsets] for e in universe], dtype=bool)
    return universe, member.reshape(len(universe), len(list_of_subsets))


class exact_cover:
    """ising/exact_cover.py: H = sum_e (1 - sum_{i: e in S_i} x_i)^2, x_i = (Z_i + 1)/2."""

    @staticmethod
    def get_operator(list_of_subsets):
        n = len(list_of_subsets)
        _, member = _membership(list_of_subsets)
        terms = _ZTerms(n)
        shift = 0
        for row in member:                       # one block per element of the universe
            idx = np.flatnonzero(row)
            Y = 1 - 0.5 * len(idx)
            a, b = np.meshgrid(idx, idx, indexing='ij')
            off = a != b
            terms.add(_bit(a[off]) | _bit(b[off]), 0.25)
            terms.add(_bit(idx), -Y)
            shift += Y * Y + 0.25 * len(idx)
        return terms.operator(), shift

    @staticmethod
    def get_solution(x):
        return 1 - x

    @staticmethod
    def check_solution_satisfiability(sol, list_of_subsets):
        universe, member = _membership(list_of_subsets)
        chosen = np.asarray(sol)[:len(list_of_subsets)] == 1
        cover = member[:, chosen].sum(axis=1) if len(universe) else np.zeros(0)
        # every element covered, and by exactly one chosen subset (no two chosen subsets overlap)
        return bool(np.all(cover == 1))

    # elements repeated inside one subset count once, as with the reference's set() intersections


class set_packing:
    """ising/set_packing.py: H = A sum_{S_i, S_j overlap} x_i x_j - sum_i x_i, A = 10."""

    @staticmethod
    def _overlap(list_of_subsets):
        _, member = _membership(list_of_subsets)
        m = member.astype(np.int64)
        return (m.T @ m) > 0

    @staticmethod
    def get_operator(list_of_subsets):
        n = len(list_of_subsets)
        A = 10
        terms = _ZTerms(n)
        i, j = _lower_pairs(n, set_packing._overlap(list_of_subsets))
        terms.add(*_pair_with_singles(i, j, A * 0.25, A * 0.25, A * 0.25))
        terms.add(_bit(np.arange(n)), -0.5)
        return terms.operator(), A * 0.25 * len(i) - 0.5 * n

    @staticmethod
    def get_solution(x):
        return 1 - x

    @staticmethod
    def check_disjoint(sol, list_of_subsets):
        chosen = np.asarray(sol)[:len(list_of_subsets)] == 1
        clash = np.tril(set_packing._overlap(list_of_subsets), -1) & np.outer(chosen, chosen)
        return not bool(clash.any())


class partition:
    """ising/partition.py: H = sum_{i>j} 2 v_i v_j Z_i Z_j, shift = sum v^2."""

    @staticmethod
    def get_operator(values):
        v = np.asarray(values)
        n = len(v)
        terms = _ZTerms(n)
        i, j = _lower_pairs(n)
        terms.add(_bit(i) | _bit(j), 2. * v[i] * v[j])
        return terms.operator(), sum(v * v)

    @staticmethod
    def partition_value(x, number_list):
        x = np.asarray(x)
        diff = np.sum(number_list[x == 0]) - np.sum(number_list[x == 1])
        return diff * diff


TspData = namedtuple('TspData', 'name dim coord w')


class tsp:
    """ising/tsp.py: n^2 qubits, qubit i*n + p = "city i is visited at step p"."""

    TspData = TspData

    @staticmethod
    def calc_distance(coord, name='tmp'):
        coord = np.asarray(coord)
        assert coord.shape[1] == 2
        delta = coord[:, None, :] - coord[None, :, :]
        w = np.rint(np.hypot(delta[..., 0], delta[..., 1]))
        np.fill_diagonal(w, 0)
        return TspData(name=name, dim=coord.shape[0], coord=coord, w=w)

    @staticmethod
    def random_tsp(n, low=0, high=100, savefile=None, seed=None, name='tmp'):
        assert n > 0
        if seed:
            aqua_globals.random_seed = seed
        coord = aqua_globals.random.uniform(low, high, (n, 2))
        ins = tsp.calc_distance(coord, name)
        if savefile:
            with open(savefile, 'w', encoding='utf8') as f:
                f.write('NAME : {}\nCOMMENT : random data\nTYPE : TSP\n'.format(ins.name))
                f.write('DIMENSION : {}\nEDGE_WEIGHT_TYPE : EUC_2D\nNODE_COORD_SECTION\n'.format(ins.dim))
                for i in range(ins.dim):
                    f.write('{} {:.4f} {:.4f}\n'.format(i + 1, ins.coord[i][0], ins.coord[i][1]))
        return ins

    @staticmethod
    def parse_tsplib_format(filename):
        """TSPLIB file with EUC_2D node coordinates (tsp.py:94-129)."""
        name, coord, in_coords = '', np.zeros((0, 2)), False
        with open(filename, encoding='utf8') as f:
            for line in f:
                if in_coords:
                    v = line.split()
                    if len(v) >= 3:
                        coord[int(v[0]) - 1] = (float(v[1]), float(v[2]))
                elif line.startswith('NAME'):
                    name = line.split(':')[1]
                elif line.startswith('DIMENSION'):
                    coord = np.zeros((int(line.split(':')[1]), 2))
                elif line.startswith('NODE_COORD_SECTION'):
                    in_coords = True
        return tsp.calc_distance(coord, name)

    @staticmethod
    def get_operator(ins, penalty=1e5):
        n = ins.dim
        w = np.asarray(ins.w, dtype=float)
        terms = _ZTerms(n * n)
        # distance part: city i at step p followed by city j at step p+1 (tsp.py:146-162)
        i, j, p = np.indices((n, n, n)).reshape(3, -1)
        sel = i != j
        i, j, p = i[sel], j[sel], p[sel]
        qa, qb = _bit(i * n + p), _bit(j * n + (p + 1) % n)
        d = w[i, j] / 4
        terms.add(np.stack([qa, qb, qa | qb], axis=1), np.stack([-d, -d, d], axis=1))
        shift = float(np.sum(d))
        # one-hot penalties: every (city, step) once; pairs of cities per step; pairs of steps per city
        terms.add(_bit(np.arange(n * n)), penalty)
        a, b = _lower_pairs(n)
        step = np.repeat(np.arange(n), len(a))
        ca, cb = np.tile(a, n), np.tile(b, n)
        qa, qb = _bit(ca * n + step), _bit(cb * n + step)
        half = np.array([-penalty / 2, -penalty / 2, penalty / 2])
        terms.add(np.stack([qa, qb, qa | qb], axis=1), half)
        city = np.repeat(np.arange(n), len(a))
        qa, qb = _bit(city * n + np.tile(a, n)), _bit(city * n + np.tile(b, n))
        terms.add(np.stack([qa, qb, qa | qb], axis=1), half)
        shift += -penalty * n * n + 2 * (penalty / 2) * n * len(a) + 2 * penalty * n
        return terms.operator(), shift

    @staticmethod
    def tsp_value(z, w):
        ret = 0.0
        for k in range(len(z) - 1):
            ret += w[z[k], z[k + 1]]
        ret += w[z[-1], z[0]]
        return ret

    @staticmethod
    def tsp_feasible(x):
        n = int(np.sqrt(len(x)))
        y = np.asarray(x)[:n * n].reshape(n, n)
        return bool(np.all(y.sum(axis=0) == 1) and np.all(y.sum(axis=1) == 1))

    @staticmethod
    def get_tsp_solution(x):
        """Cities per step; a step with != 1 city is reported as a list (tsp.py:236-258)."""
        n = int(np.sqrt(len(x)))
        y = np.asarray(x)[:n * n].reshape(n, n) >= 0.999
        z = []
        for p in range(n):
            cities = [int(c) for c in np.flatnonzero(y[:, p])]
            if len(cities) == 1:
                z.extend(cities)
            else:
                z.append(cities)
        return z


class knapsack:
    """ising/knapsack.py: C = M (W_max - sum x_i w_i - S)^2 - sum x_i v_i with S in binary on
    extra qubits, M = 10 sum(values)."""

    @staticmethod
    def get_operator(values, weights, max_weight):
        if len(values) != len(weights):
            raise ValueError("The values and weights must have the same length")
        if any(v < 0 for v in values) or any(w < 0 for w in weights):
            raise ValueError("The values and weights cannot be negative")
        if all(v == 0 for v in values):
            raise ValueError("The values cannot all be 0")
        if max_weight < 0:
            raise ValueError("max_weight cannot be negative")
        y_size = int(math.log(max_weight, 2)) + 1 if max_weight > 0 else 1
        n = len(values)
        M = float(10 * np.sum(values))
        v = np.asarray(values, dtype=float)
        u = np.concatenate([np.asarray(weights, dtype=float), 2.0 ** np.arange(y_size)])  # x then y bits
        terms = _ZTerms(n + y_size)
        shift = 0.0

        def square_block(lo, hi, scale):
            # all ordered (i, j) of one register, i == j included: Z_j, Z_i, Z_i Z_j (identity when
            # i == j: the two flips cancel, knapsack.py:217-223)
            nonlocal shift
            i, j = np.indices((hi - lo, hi - lo)).reshape(2, -1) + lo
            c = -scale * u[i] * u[j] * M
            terms.add(np.stack([_bit(j), _bit(i), _bit(i) ^ _bit(j)], axis=1), np.stack([c, c, -c], axis=1))
            shift -= float(np.sum(c))

        square_block(0, n, 0.25)
        square_block(n, n + y_size, 0.25)
        lin = max_weight * u * M
        terms.add(_bit(np.arange(n + y_size)), lin)
        shift -= float(np.sum(lin))
        i, j = np.indices((n, y_size)).reshape(2, -1)
        j = j + n
        c = -0.5 * u[i] * u[j] * M
        terms.add(np.stack([_bit(j), _bit(i), _bit(i) | _bit(j)], axis=1), np.stack([c, c, -c], axis=1))
        shift -= float(np.sum(c))
        terms.add(_bit(np.arange(n)), 0.5 * v)
        shift -= float(np.sum(0.5 * v))
        return terms.operator(), shift

    @staticmethod
    def get_solution(x, values):
        return x[:len(values)]

    @staticmethod
    def knapsack_value_weight(solution, values, weights):
        return np.sum(solution * values), np.sum(solution * weights)


class vehicle_routing:
    """ising/vehicle_routing.py: n nodes (node 0 = depot), K vehicles, one qubit per directed
    edge (i -> j), edge index i*(n-1) + (j if j < i else j - 1)."""

    @staticmethod
    def get_vehiclerouting_matrices(instance, n, K):  # noqa: N803
        instance = np.asarray(instance, dtype=float)
        N = n * (n - 1)
        A = np.max(instance) * 100
        # edge lengths: the positive entries of the matrix in row-major order (zero-length edges
        # shift the rest forward, as in the reference, vehicle_routing.py:46-51)
        positive = instance.reshape(-1)[instance.reshape(-1) > 0]
        w = np.zeros(N)
        w[np.arange(len(positive))] = positive
        source = np.repeat(np.arange(n), n - 1)
        t = np.tile(np.arange(n - 1), n)
        target = t + (t >= source)
        v = (target[None, :] == np.arange(n)[:, None]).astype(float)      # v[i, e] = edge e ends in i
        out_of = (source[None, :] == np.arange(n)[:, None]).astype(float)  # edge e starts in i
        Q = A * (out_of.T @ out_of + v.T @ v)
        not_depot = (np.arange(n) != 0).astype(float)
        g = w - 2 * A * (not_depot @ out_of + not_depot @ v) \
            - 2 * A * K * (out_of[0] + v[0])
        c = 2 * A * (n - 1) + 2 * A * (K ** 2)
        return Q, g, c

    @staticmethod
    def get_vehiclerouting_cost(instance, n, K, x_sol):  # noqa: N803
        Q, g, c = vehicle_routing.get_vehiclerouting_matrices(instance, n, K)
        x = np.around(x_sol)
        return x @ Q @ x + g @ x + c

    @staticmethod
    def get_operator(instance, n, K):  # noqa: N803
        """Binary program -> Z basis (x = (1 - Z)/2); returns the operator only, the constant is
        kept as an identity term (vehicle_routing.py:113-154)."""
        Q, g, c = vehicle_routing.get_vehiclerouting_matrices(instance, n, K)
        N = n * (n - 1)
        ones = np.ones(N)
        q_z = Q / 4
        g_z = -g / 2 - ones @ q_z - q_z @ ones
        c_z = c + (g / 2) @ ones + ones @ q_z @ ones + np.trace(q_z)
        q_z = q_z - np.diag(np.diag(q_z))
        terms = _ZTerms(N)
        k = np.flatnonzero(g_z != 0)
        terms.add(_bit(k), g_z[k])
        i, j = _lower_pairs(N, q_z != 0)
        terms.add(_bit(i) | _bit(j), 2 * q_z[i, j])
        terms.add(np.zeros(1, np.uint64), c_z)
        return terms.operator()

    @staticmethod
    def get_vehiclerouting_solution(instance, n, K, result):  # noqa: N803
        """Bit string (edge 0 first) of the largest eigenstate component."""
        del instance, K
        v = result.eigenstate
        if hasattr(v, 'device_vector'):
            v = v.device_vector()
        if hasattr(v, 'is_cuda'):
            k = int(v.abs().argmax().item())
        else:
            v = np.asarray(v)
            k = int(np.flatnonzero(v == v.max())[0])
        return (k >> np.arange(n * (n - 1))) & 1
