"""Ansatz -> psi on the device ("next" row 1).

A tiny circuit description (RY/RZ/RX/H/X/CX/CZ) executed gate by gate on a complex128 CUDA
tensor through libpsum.so, so a VQE-style energy evaluation  E(theta) = <psi(theta)|H|psi(theta)>
never moves 2^n amplitudes over PCIe.  Stands in for the simulator call inside
CircuitSampler (reference converters/circuit_sampler.py:240-342, :330-333) and the
RealAmplitudes / EfficientSU2 ansatz circuits used by the reference's VQE tests
(test/aqua/test_vqe.py:34-55).  Qubit q <-> bit q of the basis index (terra ordering).
"""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import check, lib
from .engine import _stream_ptr


def _u(name, theta=0.0):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    if name == 'ry':
        return [[c, -s], [s, c]]
    if name == 'rx':
        return [[c, -1j * s], [-1j * s, c]]
    if name == 'rz':
        return [[complex(c, -s), 0], [0, complex(c, s)]]
    if name == 'h':
        r = 1 / math.sqrt(2)
        return [[r, r], [r, -r]]
    if name == 'x':
        return [[0, 1], [1, 0]]
    raise ValueError('unknown gate %s' % name)


class StatevectorCircuit:
    """Ordered gate list; parameters are either numbers or integer indices into a vector."""

    def __init__(self, num_qubits):
        self.num_qubits = int(num_qubits)
        self.gates = []          # (name, qubits, value or ('p', index))
        self.num_parameters = 0

    def _add(self, name, qubits, value=None):
        self.gates.append((name, tuple(qubits), value))
        return self

    def new_parameter(self):
        self.num_parameters += 1
        return ('p', self.num_parameters - 1)

    def ry(self, theta, q):
        return self._add('ry', [q], theta)

    def rx(self, theta, q):
        return self._add('rx', [q], theta)

    def rz(self, theta, q):
        return self._add('rz', [q], theta)

    def h(self, q):
        return self._add('h', [q])

    def x(self, q):
        return self._add('x', [q])

    def cx(self, c, t):
        return self._add('cx', [c, t])

    def cz(self, c, t):
        return self._add('cz', [c, t])

    # -- execution ------------------------------------------------------------------------------------
    def _bound_gates(self, params):
        """[(name, qubits, 2x2 matrix or None)] with the parameters bound."""
        if params is None:
            params = []
        if len(params) != self.num_parameters:
            raise ValueError('expected %d parameters, got %d' % (self.num_parameters, len(params)))
        out = []
        for name, qs, val in self.gates:
            if name in ('cx', 'cz'):
                out.append((name, qs, None))
                continue
            theta = 0.0 if val is None else (params[val[1]] if isinstance(val, tuple) else val)
            out.append((name, qs, np.asarray(_u(name, float(theta)), dtype=np.complex128)))
        return out

    def run_unfused(self, params=None, out=None):
        """Gate by gate: one sweep of the state per gate (psum_gate_1q / psum_gate_2q)."""
        n = self.num_qubits
        if out is None:
            out = torch.empty(1 << n, dtype=torch.complex128, device='cuda')
        L = lib()
        ptr, st = C.c_void_p(out.data_ptr()), _stream_ptr()
        check(L.psum_state_basis(ptr, n, 0, st))
        buf = (C.c_double * 8)()
        for name, qs, u in self._bound_gates(params):
            if u is None:
                check(L.psum_gate_2q(ptr, n, 0 if name == 'cx' else 1, qs[0], qs[1], st))
                continue
            u = u.reshape(4)
            for i in range(4):
                buf[2 * i], buf[2 * i + 1] = u[i].real, u[i].imag
            check(L.psum_gate_1q(ptr, n, qs[0], buf, st))
        return out

    def run(self, params=None, out=None, max_block_qubits=12):
        """Executes on |0...0> and returns the complex128 CUDA state tensor.  The gate list is cut into
        blocks whose qubits fit one shared-memory tile (fuse_schedule); each block costs ONE sweep of
        the state (psum_gate_block) instead of one per gate."""
        n = self.num_qubits
        if out is None:
            out = torch.empty(1 << n, dtype=torch.complex128, device='cuda')
        L = lib()
        ptr, st = C.c_void_p(out.data_ptr()), _stream_ptr()
        check(L.psum_state_basis(ptr, n, 0, st))
        blocks = fuse_schedule(n, self._bound_gates(params), max_block_qubits)
        self.last_block_count = len(blocks)
        for free, gates in blocks:
            pos = {q: j for j, q in enumerate(free)}
            recs = np.zeros(len(gates), dtype=GATE_DT)
            for k, (name, qs, u) in enumerate(gates):
                if u is None:
                    recs[k]['kind'] = 1 if name == 'cx' else 2
                    recs[k]['b'], recs[k]['a'] = pos[qs[0]], pos[qs[1]]      # (control, target)
                else:
                    recs[k]['kind'] = 0
                    recs[k]['a'] = pos[qs[0]]
                    recs[k]['u'] = u.reshape(4).view(np.float64)
            fq = (C.c_int32 * len(free))(*free)
            check(L.psum_gate_block(ptr, n, fq, len(free), recs.ctypes.data_as(C.c_void_p), len(gates), st))
        return out


GATE_DT = np.dtype([('kind', '<i4'), ('a', '<i4'), ('b', '<i4'), ('pad', '<i4'), ('u', '<f8', (8,))])
assert GATE_DT.itemsize == 80


def fuse_schedule(n, gates, max_block_qubits=12):
    """Cuts a gate list [(name, qubits, matrix)] into blocks [(free_qubits, gates)].

    Consecutive one-qubit gates on a qubit are first merged into one 2x2 matrix.  A block then
    collects, in program order, every gate whose qubits still fit the block's free set (at most
    max_block_qubits qubits, always containing the three lowest ones so that the tile is read in
    128-byte runs); a gate that does not fit blocks its qubits, and every later gate touching a blocked
    qubit waits for a later block -- gates on disjoint qubits commute, so program order is preserved
    wherever it matters."""
    merged = []
    last_1q = {}                      # qubit -> index in merged of a trailing one-qubit gate
    for name, qs, u in gates:
        if u is not None:
            q = qs[0]
            if q in last_1q:
                k = last_1q[q]
                merged[k] = ('u', (q,), u @ merged[k][2])
            else:
                last_1q[q] = len(merged)
                merged.append(('u', (q,), u))
        else:
            for q in qs:
                last_1q.pop(q, None)
            merged.append((name, tuple(qs), None))
    low = list(range(min(3, n)))
    L = min(max(max_block_qubits, 5), 12, n)      # 3 fixed low qubits + room for any two-qubit gate
    blocks = []
    pending = merged
    while pending:
        free, blocked, block, rest = set(low), set(), [], []
        for g in pending:
            qs = set(g[1])
            if qs & blocked:
                blocked |= qs
                rest.append(g)
                continue
            need = qs - free
            if len(free) + len(need) <= L:
                free |= need
                block.append(g)
            else:
                blocked |= qs
                rest.append(g)
        for q in range(n):              # pad the tile with the lowest unused qubits
            if len(free) >= L:
                break
            free.add(q)
        if not block:
            raise RuntimeError('fuse_schedule made no progress')       # cannot happen for L >= min(5, n)
        blocks.append((sorted(free), block))
        pending = rest
    return blocks


def _entangler_pairs(n, entanglement):
    if entanglement == 'linear':
        return [(i, i + 1) for i in range(n - 1)]
    if entanglement == 'full':
        return [(i, j) for i in range(n) for j in range(i + 1, n)]
    if entanglement == 'circular':
        return ([(n - 1, 0)] if n > 2 else []) + [(i, i + 1) for i in range(n - 1)]
    raise ValueError('unknown entanglement %r' % entanglement)


class RealAmplitudes(StatevectorCircuit):
    """RY layers + CX entanglers (terra circuit library RealAmplitudes, parameter order
    theta[layer * n + qubit])."""

    def __init__(self, num_qubits, reps=3, entanglement='full'):
        super().__init__(num_qubits)
        for layer in range(reps + 1):
            for q in range(num_qubits):
                self.ry(self.new_parameter(), q)
            if layer < reps:
                for c, t in _entangler_pairs(num_qubits, entanglement):
                    self.cx(c, t)


class EfficientSU2(StatevectorCircuit):
    """RY + RZ layers + CX entanglers (parameter order: all RY of a layer, then all RZ)."""

    def __init__(self, num_qubits, reps=3, entanglement='full'):
        super().__init__(num_qubits)
        for layer in range(reps + 1):
            for q in range(num_qubits):
                self.ry(self.new_parameter(), q)
            for q in range(num_qubits):
                self.rz(self.new_parameter(), q)
            if layer < reps:
                for c, t in _entangler_pairs(num_qubits, entanglement):
                    self.cx(c, t)


class CircuitStateFn:
    """state_fns/circuit_state_fn.py stand-in: a state defined by a bound circuit; evaluating it
    prepares psi on the device."""

    def __init__(self, circuit, params=None, coeff=1.0):
        self.circuit, self.params, self.coeff = circuit, params, coeff
        self.is_measurement = False

    @property
    def num_qubits(self):
        return self.circuit.num_qubits

    def bind_parameters(self, params):
        return CircuitStateFn(self.circuit, params, self.coeff)

    def to_vector_state_fn(self):
        from .operators import VectorStateFn
        return VectorStateFn(self.circuit.run(self.params), coeff=self.coeff)

    def device_vector(self):
        return self.circuit.run(self.params)

    def eval(self, front=None):
        return self.to_vector_state_fn()


def energy_evaluation(operator, ansatz, parameters):
    """VQE._energy_evaluation (reference algorithms/minimum_eigen_solvers/vqe.py:496-541) for a
    statevector backend: parameters may hold several concatenated parameter sets
    (max_evals_grouped, vqe.py:518); returns the real means, one per set, computed entirely on
    the device."""
    p = np.asarray(parameters, dtype=float).reshape(-1, ansatz.num_parameters)
    eng = operator.engine if hasattr(operator, 'engine') else operator._engine_for()
    out = []
    buf = None
    for row in p:
        buf = ansatz.run(row, out=buf)
        out.append(eng.expect(buf).real)
    return out[0] if len(out) == 1 else np.array(out)


class VQE:
    """Thin caller around the device path (the optimiser loop itself is host-side scipy, as the
    reference's is host-side Python): minimises energy_evaluation over the ansatz parameters.
    Mirrors the result fields of reference algorithms/minimum_eigen_solvers/vqe.py:403-467
    (eigenvalue, optimal_point, optimal_value, cost_function_evals, eigenstate)."""

    def __init__(self, operator, var_form, optimizer='SLSQP', initial_point=None, maxiter=300, seed=0,
                 callback=None):
        self.operator, self.var_form = operator, var_form
        self.optimizer, self.maxiter = optimizer, maxiter
        self.initial_point = initial_point
        self.seed = seed
        self.callback = callback
        self._evals = 0

    def _cost(self, theta):
        self._evals += 1
        val = float(energy_evaluation(self.operator, self.var_form, theta))
        if self.callback is not None:
            self.callback(self._evals, theta, val, 0.0)
        return val

    def run(self):
        from scipy.optimize import minimize
        rng = np.random.default_rng(self.seed)
        x0 = np.asarray(self.initial_point, dtype=float) if self.initial_point is not None \
            else rng.uniform(-np.pi, np.pi, self.var_form.num_parameters)
        self._evals = 0
        res = minimize(self._cost, x0, method=self.optimizer, options={'maxiter': self.maxiter})
        state = self.var_form.run(res.x)
        return {'eigenvalue': complex(res.fun), 'optimal_value': float(res.fun), 'optimal_point': res.x,
                'cost_function_evals': self._evals, 'eigenstate': state, 'optimizer_message': str(res.message)}

    compute_minimum_eigenvalue = run
