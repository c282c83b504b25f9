"""`MatrixOperator` and `op_converter` of the legacy operator stack, re-hosted on the matrix-free
engine.

  qiskit/aqua/operators/legacy/matrix_operator.py:39-61, 155-205, 260-276, 319-355, 387-394
  qiskit/aqua/operators/legacy/op_converter.py:35-134

In the reference a MatrixOperator IS a scipy CSR matrix, and converting a Pauli sum to it is the
O(T * 2^n * G_x) assembly that dominates `evaluate_with_statevector`.  Here a MatrixOperator keeps
the Pauli sum: `to_matrix_operator` costs nothing, `evaluate_with_statevector` / `evolve` run on the
device kernels, and the CSR / dense views are produced on demand (column by column on the device)
for the small operators whose matrices the reference's tests inspect.  A MatrixOperator made from
a raw matrix is decomposed once into its Pauli sum with one Walsh-Hadamard transform per x-mask on
the device (O(4^n n) instead of the reference's 4^n sparse products, op_converter.py:35-39).
"""
import copy as _copy

import numpy as np
import scipy.sparse as sp

from .aqua_globals import aqua_globals
from .operators import AquaError, LegacyBaseOperator, Pauli, WeightedPauliOperator

_DECOMPOSE_MAX_QUBITS = 12


def pauli_decompose(matrix, atol=0.0, diagonal_only=False):
    """matrix (2^n x 2^n, dense or scipy sparse) -> [[weight, Pauli], ...] with
    weight = tr(P M) / 2^n, in the reference's enumeration order (labels over 'IXYZ', leftmost
    character slowest) and with its filter `weight != 0 and |weight| > atol`.

    For every x the generalised diagonal d_x[i] = M[i, i^x] holds the terms with that x-mask, and
    their weights are its Walsh-Hadamard transform over z:
        c(x, z) = (+i)^{popc(x&z)} / 2^n * sum_i (-1)^{popc(i&z)} M[i, i^x].
    """
    m = matrix.toarray() if sp.issparse(matrix) else np.asarray(matrix)
    dim = m.shape[0]
    n = int(np.log2(dim))
    if m.shape != (dim, dim) or (1 << n) != dim:
        raise AquaError('a MatrixOperator needs a 2^n x 2^n matrix')
    if n > _DECOMPOSE_MAX_QUBITS and not aqua_globals.massive:
        raise ValueError("decomposing a {}-qubit matrix into 4^{} Pauli weights is exponentially large. "
                         "Set aqua_globals.massive=True if you want to proceed.".format(n, n))
    # one Walsh-Hadamard transform per x-mask on the device (psum_pauli_decompose)
    from .engine import pauli_decompose_device
    w = pauli_decompose_device(m.astype(np.complex128)).cpu().numpy()
    xs = np.arange(dim)
    if diagonal_only:
        xs, w = xs[:1], w[:1]
    x_idx, z_idx = np.nonzero((w != 0.0) & (np.abs(w) > atol))
    x_mask = xs[x_idx]
    ny = np.array([bin(int(a) & int(b)).count('1') for a, b in zip(x_mask, z_idx)], dtype=np.int64)
    weights = w[x_idx, z_idx] * np.array([1, 1j, -1, -1j])[ny & 3]
    # enumeration rank of the label: per qubit I < X < Y < Z, qubit n-1 most significant
    rank = np.zeros(len(x_mask), dtype=np.int64)
    for q in range(n):
        xb, zb = (x_mask >> q) & 1, (z_idx >> q) & 1
        rank += np.where(xb == 1, np.where(zb == 1, 2, 1), np.where(zb == 1, 3, 0)) << (2 * q)
    order = np.argsort(rank, kind='stable')
    return [[complex(weights[k]), Pauli.from_masks(int(x_mask[k]), int(z_idx[k]), n)] for k in order]


class MatrixOperator(LegacyBaseOperator):
    """Operator given by its matrix; held as a Pauli sum (see module docstring)."""

    def __init__(self, matrix, basis=None, z2_symmetries=None, atol=1e-12, name=None):
        self._basis = basis
        self._z2_symmetries = z2_symmetries
        self._name = name
        self._atol = atol
        if matrix is None:
            self._sum = None
        elif isinstance(matrix, WeightedPauliOperator):
            self._sum = matrix
        else:
            nnz = matrix.nnz if sp.issparse(matrix) else np.count_nonzero(matrix)
            if nnz == 0:
                self._sum = None
            else:
                dense = matrix.toarray() if sp.issparse(matrix) else np.asarray(matrix)
                diagonal = np.count_nonzero(dense) == np.count_nonzero(np.diag(dense))
                self._sum = WeightedPauliOperator(paulis=pauli_decompose(dense, 0.0, diagonal_only=diagonal))
                if self._sum.is_empty():
                    self._sum = None

    # -- bookkeeping ----------------------------------------------------------------------------
    name = property(lambda self: self._name, lambda self, v: setattr(self, '_name', v))
    basis = property(lambda self: self._basis)
    z2_symmetries = property(lambda self: self._z2_symmetries)

    @property
    def atol(self):
        return self._atol

    @atol.setter
    def atol(self, new_value):
        self._atol = new_value

    @property
    def pauli_sum(self):
        """The WeightedPauliOperator this operator is stored as (None when empty)."""
        return self._sum

    def is_empty(self):
        return self._sum is None or self._sum.is_empty()

    @property
    def num_qubits(self):
        return 0 if self.is_empty() else self._sum.num_qubits

    def copy(self):
        return _copy.deepcopy(self)

    def __str__(self):
        length = '{}x{}'.format(2 ** self.num_qubits, 2 ** self.num_qubits)
        return 'Representation: {}, qubits: {}, size: {}'.format('matrix', self.num_qubits, length)

    def print_details(self):
        return str(self._matrix)

    # -- matrix views (device-produced; small operators) ----------------------------------------
    @property
    def dense_matrix(self):
        if self.is_empty():
            raise AquaError('Operator is empty, check the operator.')
        return self._sum.to_opflow().to_matrix()

    @property
    def _matrix(self):
        return None if self.is_empty() else sp.csr_matrix(self.dense_matrix)

    @property
    def dia_matrix(self):
        """The diagonal as a vector when the operator is diagonal, else None (:155-161)."""
        if self.is_empty():
            return None
        if any(p.x_mask for _, p in self._sum.paulis):
            return None
        import torch
        ones = torch.ones(1 << self.num_qubits, dtype=torch.complex128, device='cuda')
        return self._sum.engine.apply(ones).cpu().numpy()        # y = H.1 is the (complex) diagonal

    @property
    def matrix(self):
        dia = self.dia_matrix
        return self._matrix if dia is None else dia

    def to_opflow(self):
        return self._sum.to_opflow()

    # -- algebra (on the Pauli sums; the same linear maps as the reference's CSR arithmetic) -----
    def _binary(self, other, fn):
        a = None if self.is_empty() else self._sum
        b = None if other.is_empty() else other._sum
        return fn(a, b)

    def add(self, other, copy=False):
        out = self.copy() if copy else self
        out._sum = self._binary(other, lambda a, b: b if a is None else (a if b is None else a + b))
        return out

    def sub(self, other, copy=False):
        out = self.copy() if copy else self
        out._sum = self._binary(other, lambda a, b: (None if b is None else -b) if a is None
                                else (a if b is None else a - b))
        return out

    def __add__(self, other):
        return self.add(other, copy=True)

    def __iadd__(self, other):
        return self.add(other, copy=False)

    def __sub__(self, other):
        return self.sub(other, copy=True)

    def __isub__(self, other):
        return self.sub(other, copy=False)

    def __neg__(self):
        out = self.copy()
        if not out.is_empty():
            out._sum = -out._sum
        return out

    def __mul__(self, other):
        if self.is_empty() or other.is_empty():
            return MatrixOperator(None)
        return MatrixOperator(self._sum * other._sum)

    def __eq__(self, other):
        if not isinstance(other, MatrixOperator):
            return NotImplemented
        if self.is_empty() or other.is_empty():
            return self.is_empty() and other.is_empty()
        a = {(p.x_mask, p.z_mask): complex(w) for w, p in self._sum.paulis}
        b = {(p.x_mask, p.z_mask): complex(w) for w, p in other._sum.paulis}
        return self.num_qubits == other.num_qubits and a == b

    __hash__ = None

    def chop(self, threshold=None, copy=False):
        """Zero the real / imaginary parts of MATRIX ELEMENTS below the threshold (:126-153); the
        matrix is materialised for this, so small operators only."""
        threshold = self._atol if threshold is None else threshold
        op = self.copy() if copy else self
        if op.is_empty():
            return op
        m = op.dense_matrix
        re = np.where(np.abs(m.real) >= threshold, m.real, 0.0)
        im = np.where(np.abs(m.imag) >= threshold, m.imag, 0.0)
        chopped = MatrixOperator(re + 1j * im)
        op._sum = chopped._sum
        return op

    # -- the hot path ---------------------------------------------------------------------------
    def evaluate_with_statevector(self, quantum_state):
        """(<psi|M|psi>, 0.0) (:260-276) through the device kernels."""
        if self.is_empty():
            raise AquaError("Operator is empty, check the operator.")
        return self._sum.evaluate_with_statevector(quantum_state)

    def evolve(self, state_in, evo_time=0, num_time_slices=0, expansion_mode='trotter', expansion_order=1):
        """exp(-i evo_time M) state_in; only the exact branch (num_time_slices = 0, :354-355)."""
        if self.is_empty():
            raise AquaError("Operator is empty, check the operator.")
        del expansion_mode, expansion_order
        return self._sum.evolve(state_in, evo_time, num_time_slices)


class op_converter:  # namespace mirroring legacy/op_converter.py
    @staticmethod
    def to_matrix_operator(operator):
        """No matrix is assembled: the MatrixOperator shares the Pauli sum (op_converter.py:103-134)."""
        if operator.__class__ == WeightedPauliOperator:
            if operator.is_empty():
                return MatrixOperator(None)
            return MatrixOperator(operator, z2_symmetries=operator.z2_symmetries, name=operator.name)
        if operator.__class__ == MatrixOperator:
            return operator
        raise AquaError("Unsupported type to convert to MatrixOperator: {}".format(operator.__class__))

    @staticmethod
    def to_weighted_pauli_operator(operator):
        """Pauli weights tr(P M)/2^n above the operator's atol (op_converter.py:42-100)."""
        if operator.__class__ == WeightedPauliOperator:
            return operator
        if operator.__class__ == MatrixOperator:
            if operator.is_empty():
                return WeightedPauliOperator(paulis=[])
            paulis = [[w, p] for w, p in _enumeration_order(operator.pauli_sum)
                      if w != 0.0 and np.abs(w) > operator.atol]
            return WeightedPauliOperator(paulis, z2_symmetries=operator.z2_symmetries, name=operator.name)
        raise AquaError("Unsupported type to convert to WeightedPauliOperator: {}".format(operator.__class__))


def _enumeration_order(pauli_sum):
    """Terms sorted the way the reference enumerates labels: 'IXYZ' product, leftmost slowest."""
    digit = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}
    n = pauli_sum.num_qubits

    def rank(p):
        return sum(digit[((p.x_mask >> q) & 1, (p.z_mask >> q) & 1)] << (2 * q) for q in range(n))

    return sorted(((complex(w), p) for w, p in pauli_sum.paulis), key=lambda t: rank(t[1]))
