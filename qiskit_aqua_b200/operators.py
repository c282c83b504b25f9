"""Host-side mirror of the reference's operator interface for the hot path (same names, argument
meaning and error behaviour), sitting on PauliSumEngine / libpsum.so.

Mirrored reference pieces (qiskit-aqua 0.10.0):
  * Pauli stand-in for terra's ``qiskit.quantum_info.Pauli`` (label / (z, x) constructor, .z, .x,
    .to_label()) -- only what the hot path uses.
  * WeightedPauliOperator  qiskit/aqua/operators/legacy/weighted_pauli_operator.py:40-623
  * opflow leaves PauliOp / SummedOp / ListOp / StateFn / MatrixExpectation:
      primitive_ops/pauli_op.py, list_ops/summed_op.py, state_fns/{vector,operator}_state_fn.py,
      expectations/matrix_expectation.py:32-49
No matrix (CSR or dense) is ever built: every evaluation is a matrix-free H.psi on the device.
"""
import copy as _copy
import json
import numbers
import operator as _operator

import numpy as np

from .engine import PauliSumEngine, as_device_state

EVAL_SIG_DIGITS = 18  # reference operator_globals.py:33


class AquaError(Exception):
    """Reference qiskit/aqua/aqua_error.py:16-30."""

    def __init__(self, *message):
        super().__init__(' '.join(str(m) for m in message))
        self.message = ' '.join(str(m) for m in message)

    def __str__(self):
        return repr(self.message)


# ------------------------------------------------------------------------------------------------
# Pauli stand-in
# ------------------------------------------------------------------------------------------------
class Pauli:
    """Minimal stand-in for terra's Pauli: ``Pauli('XYZ')`` (big-endian label, leftmost char is
    the highest qubit) or ``Pauli((z, x))`` with little-endian boolean arrays."""

    __slots__ = ('_x', '_z', '_n')

    def __init__(self, data=None, x=None, z=None, label=None):
        if label is not None:
            data = label
        if isinstance(data, Pauli):
            self._x, self._z, self._n = data._x, data._z, data._n
            return
        if isinstance(data, str):
            n = len(data)
            xm = zm = 0
            for pos, ch in enumerate(data):
                bit = 1 << (n - 1 - pos)
                if ch == 'X':
                    xm |= bit
                elif ch == 'Z':
                    zm |= bit
                elif ch == 'Y':
                    xm |= bit
                    zm |= bit
                elif ch != 'I':
                    raise ValueError('Pauli string is not valid: %r' % data)
            self._x, self._z, self._n = xm, zm, n
            return
        if data is not None:
            z, x = data
        if z is None or x is None:
            raise ValueError('Pauli needs a label or (z, x) arrays')
        z = np.asarray(z).astype(bool)
        x = np.asarray(x).astype(bool)
        if z.shape != x.shape:
            raise ValueError('z and x must have the same length')
        self._n = int(z.shape[0])
        self._x = sum(1 << q for q in range(self._n) if x[q])
        self._z = sum(1 << q for q in range(self._n) if z[q])

    @classmethod
    def from_label(cls, label):
        return cls(label)

    @classmethod
    def from_masks(cls, x_mask, z_mask, num_qubits):
        p = cls.__new__(cls)
        p._x, p._z, p._n = int(x_mask), int(z_mask), int(num_qubits)
        return p

    @property
    def num_qubits(self):
        return self._n

    def __len__(self):
        return self._n

    @property
    def x(self):
        return np.array([(self._x >> q) & 1 for q in range(self._n)], dtype=bool)

    @property
    def z(self):
        return np.array([(self._z >> q) & 1 for q in range(self._n)], dtype=bool)

    @property
    def x_mask(self):
        return self._x

    @property
    def z_mask(self):
        return self._z

    def to_label(self):
        return ''.join('IXZY'[((self._x >> q) & 1) + 2 * ((self._z >> q) & 1)]
                       for q in range(self._n - 1, -1, -1))

    def __str__(self):
        return self.to_label()

    def __repr__(self):
        return "Pauli('%s')" % self.to_label()

    def __eq__(self, other):
        return isinstance(other, Pauli) and (self._x, self._z, self._n) == (other._x, other._z, other._n)

    def __hash__(self):
        return hash((self._x, self._z, self._n))

    def dot(self, other):
        """Product self.other as (Pauli, phase) with phase in {1, i, -1, -i} (label Paulis)."""
        cx, cz = self._x ^ other._x, self._z ^ other._z
        e = (bin(self._x & self._z).count('1') + bin(other._x & other._z).count('1')
             - bin(cx & cz).count('1') + 2 * bin(self._z & other._x).count('1')) % 4
        return Pauli.from_masks(cx, cz, self._n), (1, 1j, -1, -1j)[e]


def _pack(num_qubits, pairs):
    """[(coeff, Pauli)] -> engine arrays."""
    T = len(pairs)
    x = np.fromiter((p.x_mask for _, p in pairs), dtype=np.uint64, count=T)
    z = np.fromiter((p.z_mask for _, p in pairs), dtype=np.uint64, count=T)
    y = np.fromiter((bin(p.x_mask & p.z_mask).count('1') & 3 for _, p in pairs), dtype=np.uint8, count=T)
    c = np.fromiter((complex(w) for w, _ in pairs), dtype=np.complex128, count=T)
    return x, z, y, c


# ------------------------------------------------------------------------------------------------
# legacy WeightedPauliOperator
# ------------------------------------------------------------------------------------------------
class LegacyBaseOperator:
    pass


class WeightedPauliOperator(LegacyBaseOperator):
    """Drop-in for legacy/weighted_pauli_operator.py:40 (hot-path subset)."""

    def __init__(self, paulis, basis=None, z2_symmetries=None, atol=1e-12, name=None):
        self._name = name
        self._z2_symmetries = z2_symmetries
        self._paulis = [[w, p if isinstance(p, Pauli) else Pauli(p)] for w, p in paulis]
        self._basis = basis
        self._engine = None
        self.simplify()
        self._atol = atol

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_list(cls, paulis, weights=None, name=None):
        if weights is not None and len(weights) != len(paulis):
            raise ValueError("The length of weights and paulis must be the same.")
        if weights is None:
            weights = [1.0] * len(paulis)
        return cls(paulis=[[w, p] for w, p in zip(weights, paulis)], name=name)

    @classmethod
    def from_dict(cls, dictionary, before_04=False):
        if 'paulis' not in dictionary:
            raise AquaError('Dictionary missing "paulis" key')
        paulis = []
        for op in dictionary['paulis']:
            if 'label' not in op:
                raise AquaError('Dictionary missing "label" key')
            if 'coeff' not in op:
                raise AquaError('Dictionary missing "coeff" key')
            coeff = op['coeff']
            if 'real' not in coeff:
                raise AquaError('Dictionary missing "real" key')
            w = complex(coeff['real'], coeff['imag']) if 'imag' in coeff else coeff['real']
            label = op['label'][::-1] if before_04 else op['label']
            paulis.append([w, Pauli(label)])
        return cls(paulis=paulis)

    @classmethod
    def from_file(cls, file_name, before_04=False):
        with open(file_name, 'r', encoding='utf8') as f:
            return cls.from_dict(json.load(f), before_04=before_04)

    def to_dict(self):
        out = {'paulis': []}
        for coeff, pauli in self._paulis:
            entry = {'label': pauli.to_label()}
            if isinstance(coeff, complex):
                entry['coeff'] = {'real': float(np.real(coeff)), 'imag': float(np.imag(coeff))}
            else:
                entry['coeff'] = {'real': coeff}
            out['paulis'].append(entry)
        return out

    def to_file(self, file_name):
        with open(file_name, 'w', encoding='utf8') as f:
            json.dump(self.to_dict(), f)

    # -- container semantics ----------------------------------------------------------------------
    @property
    def paulis(self):
        return self._paulis

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        self._name = value

    @property
    def atol(self):
        return self._atol

    @atol.setter
    def atol(self, value):
        self._atol = value

    @property
    def num_qubits(self):
        return self._paulis[0][1].num_qubits if not self.is_empty() else 0

    @property
    def z2_symmetries(self):
        return self._z2_symmetries

    def is_empty(self):
        return not self._paulis or not self._paulis[0]

    def copy(self):
        new = _copy.copy(self)
        new._paulis = [[w, p] for w, p in self._paulis]
        new._engine = None
        return new

    def simplify(self, copy=False):
        """Merge equal Paulis in first-seen order, drop exact zeros (reference :341-400)."""
        op = self.copy() if copy else self
        table, merged = {}, []
        for w, p in op._paulis:
            slot = table.get(p)
            if slot is None:
                table[p] = len(merged)
                merged.append([w, p])
            else:
                merged[slot][0] = merged[slot][0] + w
        op._paulis = merged
        op._engine = None
        op.chop(0.0)
        return op

    def chop(self, threshold=None, copy=False):
        """Component-wise truncation of weights (reference :419-479)."""
        threshold = self._atol if threshold is None else threshold
        op = self.copy() if copy else self
        kept = []
        for w, p in op._paulis:
            re = np.real(w) if abs(np.real(w)) >= threshold else 0.0
            im = np.imag(w) if abs(np.imag(w)) >= threshold else 0.0
            if re == 0.0 and im == 0.0:
                continue
            kept.append([re + 1j * im, p])
        op._paulis = kept
        op._engine = None
        return op

    def rounding(self, decimals, copy=False):
        op = self.copy() if copy else self
        op._paulis = [[np.around(w, decimals=decimals), p] for w, p in op._paulis]
        op._engine = None
        return op

    def __eq__(self, other):
        """Equal after both sides are simplified (in place, as the reference does): same Paulis with
        exactly equal weights (reference :147-166)."""
        if not isinstance(other, WeightedPauliOperator):
            return False
        self.simplify()
        other.simplify()
        if len(self._paulis) != len(other._paulis):
            return False
        theirs = {p: w for w, p in other._paulis}
        return all(p in theirs and theirs[p] == w for w, p in self._paulis)

    __hash__ = None

    def _add_or_sub(self, other, operation, copy=True):
        """Weights of Paulis present on both sides are combined in place, the others appended in
        `other`'s order; zero results are kept until simplify()/chop() (reference :168-203)."""
        if not self.is_empty() and not other.is_empty() and self.num_qubits != other.num_qubits:
            raise AquaError('Can not add/sub two operators with different number of qubits.')
        ret = self.copy() if copy else self
        table = {p: k for k, (_, p) in enumerate(ret._paulis)}
        for w, p in other._paulis:
            k = table.get(p)
            if k is not None:
                ret._paulis[k][0] = operation(ret._paulis[k][0], w)
            else:
                table[p] = len(ret._paulis)
                ret._paulis.append([operation(0.0, w), p])
        ret._engine = None
        return ret

    def add(self, other, copy=False):
        return self._add_or_sub(other, _operator.add, copy=copy)

    def sub(self, other, copy=False):
        return self._add_or_sub(other, _operator.sub, copy=copy)

    def __add__(self, other):
        return self.add(other, copy=True)

    def __iadd__(self, other):
        return self.add(other, copy=False)

    def __sub__(self, other):
        return self.sub(other, copy=True)

    def __isub__(self, other):
        return self.sub(other, copy=False)

    def _scaling_weight(self, scaling_factor, copy=False):
        if not isinstance(scaling_factor, numbers.Number):
            raise ValueError("Type of scaling factor is a valid type. {} if given.".format(scaling_factor.__class__))
        ret = self.copy() if copy else self
        ret._paulis = [[w * scaling_factor, p] for w, p in ret._paulis]
        ret._engine = None
        return ret

    def multiply(self, other):
        """self * other with the Pauli phases tracked; products are accumulated term by term, so
        weights that cancel stay as zero entries until simplify() (reference :274-292)."""
        ret = WeightedPauliOperator(paulis=[])
        table = {}
        for w1, p1 in self._paulis:
            for w2, p2 in other._paulis:
                p, ph = p1.dot(p2)
                w = w1 * w2 * ph
                if np.real(w) == 0.0 and np.imag(w) == 0.0:
                    continue
                k = table.get(p)
                if k is not None:
                    ret._paulis[k][0] = ret._paulis[k][0] + w
                else:
                    table[p] = len(ret._paulis)
                    ret._paulis.append([0.0 + w, p])
        return ret

    def __mul__(self, other):
        if isinstance(other, numbers.Number):
            return self._scaling_weight(other, copy=True)
        if isinstance(other, WeightedPauliOperator):
            return self.multiply(other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, numbers.Number):
            return self._scaling_weight(other, copy=True)
        return NotImplemented

    def __neg__(self):
        return self._scaling_weight(-1.0, copy=True)

    def commute_with(self, other):
        """[self, other] == 0 (reference :481-483, legacy/common.py:213-227)."""
        com = self * other - other * self
        com.simplify()
        return bool(com.is_empty())

    def anticommute_with(self, other):
        com = self * other + other * self
        com.simplify()
        return bool(com.is_empty())

    @property
    def basis(self):
        """Grouping info [(basis Pauli, [indices into paulis])]; ungrouped: every Pauli its own."""
        if self._basis is None or len(self._basis) != len(self._paulis):
            return [(p, [k]) for k, (_, p) in enumerate(self._paulis)]
        return self._basis

    def reorder_paulis(self):
        """Ungrouped operator: the paulis as they are (reference :840-866)."""
        return self._paulis

    def __str__(self):
        name = '' if not self._name else '{}: '.format(self._name)
        return '{}Representation: paulis, qubits: {}, size: {}'.format(name, self.num_qubits, len(self._paulis))

    def print_details(self):
        if self.is_empty():
            return 'Operator is empty.'
        return ''.join('%s\t%s\n' % (p.to_label(), w) for w, p in self._paulis)

    # -- the hot path ---------------------------------------------------------------------------------
    @property
    def engine(self):
        """The packed operator on libpsum.so (packed once, cached until the operator mutates)."""
        if self._engine is None:
            if self.is_empty():
                raise AquaError('Operator is empty, check the operator.')
            self._engine = PauliSumEngine(self.num_qubits, *_pack(self.num_qubits, self._paulis))
        return self._engine

    def evaluate_with_statevector(self, quantum_state):
        """(<psi|H|psi>, 0.0) -- reference :605-623, without building the CSR matrix.
        ``quantum_state`` may be a numpy array (copied to the device) or a complex128 CUDA tensor."""
        if self.is_empty():
            raise AquaError('Operator is empty, check the operator.')
        if not getattr(quantum_state, 'is_cuda', False) and self.engine.world == 1:
            return self.engine.expect_host(quantum_state), 0.0      # H2D pipelined with the passes
        psi = as_device_state(quantum_state, 1 << self.num_qubits)
        return self.engine.expect(psi), 0.0

    def evaluate_with_result(self, result, statevector_mode, use_simulator_snapshot_mode=False,
                             circuit_name_prefix=''):
        """(mean, std_dev) from a backend result -- reference :757-812.  Statevector branch
        (:793-802): ``result.get_statevector(prefix + 'psi')`` is psi and
        ``result.get_statevector(prefix + label)`` is P|psi> for every non-identity Pauli; the mean is
        sum_k w_k <psi|P_k psi> (identity terms add w_k).  The T overlaps are device reductions; the
        states may be numpy arrays or CUDA tensors.  The snapshot branch (:784-792) reads the value
        the simulator computed.  The shot-based branch (:803-811) is not part of the H.psi path."""
        if self.is_empty():
            raise AquaError('Operator is empty, check the operator.')
        if use_simulator_snapshot_mode:
            avg = result.data(circuit_name_prefix + 'snapshot_mode')['snapshots']['expectation_value']['expval'][0]['value']
            if isinstance(avg, (list, tuple)):
                avg = avg[0] + 1j * avg[1]
            return avg, 0.0
        if not statevector_mode:
            raise AquaError('evaluate_with_result: only the statevector (and snapshot) branches are provided; '
                            'the shot-based estimator is outside the Pauli-sum H.psi path')
        from .engine import vec_dot
        psi = as_device_state(result.get_statevector(circuit_name_prefix + 'psi'), 1 << self.num_qubits)
        avg = 0.0
        for weight, pauli in self._paulis:
            if pauli.x_mask == 0 and pauli.z_mask == 0:
                avg += weight
            else:
                psi_i = as_device_state(result.get_statevector(circuit_name_prefix + pauli.to_label()),
                                        1 << self.num_qubits)
                avg += weight * vec_dot(psi, psi_i)
        return avg, 0.0

    def evolve(self, state_in, evo_time=0, num_time_slices=0, tol=1e-10):
        """exp(-i * evo_time * H) @ state_in -- the exact branch (num_time_slices == 0) of
        MatrixOperator.evolve (legacy/matrix_operator.py:319-355), by Krylov exponentiation on the
        device.  Returns a numpy array for numpy input, else a CUDA tensor."""
        if self.is_empty():
            raise AquaError('Operator is empty, check the operator.')
        if num_time_slices < 0 or not isinstance(num_time_slices, int):
            raise ValueError('Number of time slices should be a non-negative integer.')
        if num_time_slices != 0:
            raise ValueError('only the exact evolution (num_time_slices=0) is provided on this path')
        out, _ = self.engine.evolve(state_in, float(np.real(evo_time)), tol=tol)
        return out if hasattr(state_in, 'is_cuda') else out.cpu().numpy()

    def to_opflow(self, reverse_endianness=False):
        """SummedOp of PauliOps (reference :96-117)."""
        ops = []
        for w, p in self._paulis:
            pauli = Pauli(p.to_label()[::-1]) if reverse_endianness else p
            coeff = np.real(w) if np.isreal(w) else w
            ops.append(PauliOp(pauli, coeff=coeff))
        if not ops:
            return SummedOp([])
        return ops[0] if len(ops) == 1 else SummedOp(ops)


# ------------------------------------------------------------------------------------------------
# opflow subset
# ------------------------------------------------------------------------------------------------
class OperatorBase:
    def __add__(self, other):
        return self.add(other)

    def __radd__(self, other):
        return self if other == 0 else self.add(other)

    def __sub__(self, other):
        return self.add(-other)

    def __neg__(self):
        return self.mul(-1.0)

    def __mul__(self, other):
        return self.mul(other)

    __rmul__ = __mul__

    def __xor__(self, other):
        if isinstance(other, numbers.Integral):
            return self.tensorpower(other)
        return self.tensor(other)

    def __matmul__(self, other):
        return self.compose(other)

    def __invert__(self):
        return self.adjoint()

    def tensorpower(self, other):
        out = self
        for _ in range(int(other) - 1):
            out = out.tensor(self)
        return out


def _terms_of(op):
    """Flatten PauliOp / SummedOp trees into [(coeff, Pauli)]."""
    if isinstance(op, PauliOp):
        return [(op.coeff, op.primitive)]
    if isinstance(op, SummedOp):
        out = []
        for sub in op.oplist:
            out += [(op.coeff * w, p) for w, p in _terms_of(sub)]
        return out
    if isinstance(op, WeightedPauliOperator):
        return [(w, p) for w, p in op.paulis]
    raise TypeError('not a Pauli-sum operator: %r' % type(op))


class _PauliSumMixin(OperatorBase):
    _engine = None

    def _engine_for(self):
        if self._engine is None:
            terms = _terms_of(self)
            if not terms:
                raise AquaError('Operator is empty, check the operator.')
            self._engine = PauliSumEngine(self.num_qubits, *_pack(self.num_qubits, terms))
        return self._engine

    def to_legacy_op(self):
        return WeightedPauliOperator(paulis=[[w, p] for w, p in _terms_of(self)])

    def _check_massive(self, method, matrix, num_qubits, massive):
        """operator_base.py:618-648: refuse dense conversions above 16 qubits unless massive."""
        from .aqua_globals import aqua_globals
        if num_qubits > 16 and not massive and not aqua_globals.massive:
            dim = 2 ** num_qubits
            if matrix:
                obj_type = 'matrix'
                dimensions = '{}x{}'.format(dim, dim)
            else:
                obj_type = 'vector'
                dimensions = str(dim)
            raise ValueError(
                "'{}' will return an exponentially large {}, in this case '{}' elements. "
                "Set aqua_globals.massive=True or the method argument massive=True "
                "if you want to proceed.".format(method, obj_type, dimensions))

    def to_matrix(self, massive=False):
        """Dense matrix (primitive_ops/pauli_op.py:151-153, list_ops/list_op.py:306-318), written by one
        device kernel: H[i, i ^ x] = D_x(i) per x-group (psum_to_dense).  Meant for the small
        operators the reference's tests inspect; the hot path never builds it."""
        n = self.num_qubits
        self._check_massive('to_matrix', True, n, massive)
        return self._engine_for().to_dense().cpu().numpy()

    def to_spmatrix(self):
        """scipy CSR of the operator (pauli_op.py:164, list_op.py:320-329); small operators only."""
        import scipy.sparse as sp
        return sp.csr_matrix(self.to_matrix())

    def add(self, other):
        if other == 0:
            return self
        return SummedOp([self, other])

    def compose(self, other):
        if isinstance(other, StateFnBase):
            return ComposedOp([self, other])
        a = WeightedPauliOperator(paulis=[[w, p] for w, p in _terms_of(self)])
        b = WeightedPauliOperator(paulis=[[w, p] for w, p in _terms_of(other)])
        return (a * b).to_opflow()

    def eval(self, front=None):
        """H|front> as a VectorStateFn (matrix-free; primitive_ops/matrix_op.py:182-209)."""
        if front is None:
            return self
        if isinstance(front, ListOp) and not isinstance(front, SummedOp):
            return ListOp([self.eval(f) for f in front.oplist], coeff=front.coeff)
        if not isinstance(front, VectorStateFn):
            front = StateFn(front)
        y = self._engine_for().apply(front.device_vector())
        return VectorStateFn(y, coeff=front.coeff)


class PauliOp(_PauliSumMixin):
    """primitive_ops/pauli_op.py (matrix-free)."""

    def __init__(self, primitive, coeff=1.0):
        self.primitive = primitive if isinstance(primitive, Pauli) else Pauli(primitive)
        self.coeff = coeff

    @property
    def num_qubits(self):
        return self.primitive.num_qubits

    def mul(self, scalar):
        return PauliOp(self.primitive, self.coeff * scalar)

    def adjoint(self):
        return PauliOp(self.primitive, np.conj(self.coeff))

    def tensor(self, other):
        if isinstance(other, PauliOp):
            return PauliOp(Pauli(self.primitive.to_label() + other.primitive.to_label()),
                           self.coeff * other.coeff)
        if isinstance(other, SummedOp):
            return SummedOp([self.tensor(o) for o in other.oplist], coeff=other.coeff)
        return NotImplemented

    def __eq__(self, other):
        if isinstance(other, numbers.Number):
            return self.coeff == other == 0
        return isinstance(other, PauliOp) and self.primitive == other.primitive and self.coeff == other.coeff

    def __hash__(self):
        return hash((self.primitive, complex(self.coeff)))

    def __repr__(self):
        return 'PauliOp(%r, coeff=%r)' % (self.primitive, self.coeff)


class ListOp(OperatorBase):
    """list_ops/list_op.py -- plain container (no combination)."""

    def __init__(self, oplist, coeff=1.0):
        self.oplist = list(oplist)
        self.coeff = coeff

    distributive = True           # list_ops/list_op.py:108-118 (ComposedOp overrides it with False)

    @staticmethod
    def combo_fn(values):         # a plain ListOp keeps the list of results (list_op.py:60, lambda x: x)
        return values

    def __len__(self):
        return len(self.oplist)

    def __getitem__(self, i):
        return self.oplist[i]

    def __iter__(self):
        return iter(self.oplist)

    @property
    def num_qubits(self):
        return self.oplist[0].num_qubits if self.oplist else 0

    def mul(self, scalar):
        return self.__class__(self.oplist, self.coeff * scalar)

    def eval(self, front=None):
        return [self.coeff * op.eval(front) if front is not None else op.eval() for op in self.oplist]


class SummedOp(ListOp, _PauliSumMixin):
    """list_ops/summed_op.py: combo_fn = sum."""

    @staticmethod
    def combo_fn(values):
        return sum(values)

    def __init__(self, oplist, coeff=1.0):
        ListOp.__init__(self, oplist, coeff)
        self._engine = None

    def add(self, other):
        if other == 0:
            return self
        if isinstance(other, SummedOp) and other.coeff == self.coeff:
            return SummedOp(self.oplist + other.oplist, self.coeff)
        return SummedOp(self.oplist + [other.mul(1.0 / self.coeff) if self.coeff != 1.0 else other], self.coeff)

    def adjoint(self):
        return SummedOp([o.adjoint() for o in self.oplist], np.conj(self.coeff))

    def tensor(self, other):
        others = other.oplist if isinstance(other, SummedOp) else [other]
        oc = other.coeff if isinstance(other, SummedOp) else 1.0
        return SummedOp([a.tensor(b) for a in self.oplist for b in others], self.coeff * oc)

    def __eq__(self, other):
        if isinstance(other, numbers.Number):
            return other == 0 and (self.coeff == 0 or not self.oplist)
        return isinstance(other, SummedOp) and self.oplist == other.oplist and self.coeff == other.coeff

    __hash__ = None

    eval = _PauliSumMixin.eval


class StateFnBase(OperatorBase):
    pass


class VectorStateFn(StateFnBase):
    """state_fns/vector_state_fn.py: a state (or measurement) backed by a complex128 vector that
    stays on the device.  As in the reference, ``adjoint()`` stores the CONJUGATED vector (and
    conj(coeff)), and ``eval`` / ``to_matrix`` then use the stored vector as is (:76-79, :139-142,
    :190-192: a plain, non-conjugating ``np.dot``)."""

    def __init__(self, primitive, coeff=1.0, is_measurement=False, _conj=None):
        self._dev = as_device_state(primitive)
        self._conj = _conj            # cached conj(primitive), shared with the adjoint
        self.coeff = coeff
        self.is_measurement = is_measurement

    @property
    def num_qubits(self):
        return int(self._dev.numel()).bit_length() - 1

    @property
    def primitive(self):
        return self._dev

    def device_vector(self):
        return self._dev

    def _conj_vector(self):
        if self._conj is None:
            self._conj = self._dev.conj().resolve_conj()
        return self._conj

    def to_matrix(self, massive=False):
        v = self._dev.cpu().numpy() * self.coeff
        return v.reshape(1, -1) if self.is_measurement else v

    def mul(self, scalar):
        return VectorStateFn(self._dev, self.coeff * scalar, self.is_measurement, _conj=self._conj)

    def add(self, other):
        if not isinstance(other, VectorStateFn):
            other = StateFn(other)
        if other.is_measurement != self.is_measurement:
            raise ValueError('cannot add a state and a measurement')
        return VectorStateFn(self._dev * self.coeff + other._dev * other.coeff, 1.0, self.is_measurement)

    def adjoint(self):
        return VectorStateFn(self._conj_vector(), np.conj(self.coeff), not self.is_measurement, _conj=self._dev)

    def compose(self, other):
        return ComposedOp([self, other])

    def eval(self, front=None):
        """sum_i self_i front_i for a measurement (vector_state_fn.py:165-196, rounded to 18 digits);
        for ``~StateFn(v)`` the stored vector is conj(v), i.e. the value is <v|front>."""
        if front is None:
            return self
        if not self.is_measurement:
            raise ValueError('Cannot compute overlap with StateFn or Operator if not Measurement.')
        if isinstance(front, ListOp) and front.distributive:
            return front.combo_fn([self.eval(f.mul(front.coeff)) for f in front.oplist])
        if not isinstance(front, VectorStateFn):
            front = StateFn(front)
        from .engine import vec_dot
        # vec_dot conjugates its first argument: hand it conj(stored) to get the plain dot product
        val = vec_dot(self._conj_vector(), front.device_vector()) * self.coeff * front.coeff
        return complex(np.round(val, decimals=EVAL_SIG_DIGITS))


class OperatorStateFn(StateFnBase):
    """state_fns/operator_state_fn.py: ~StateFn(H) is the measurement <.|H|.>."""

    def __init__(self, primitive, coeff=1.0, is_measurement=False):
        self.primitive = primitive
        self.coeff = coeff
        self.is_measurement = is_measurement

    @property
    def num_qubits(self):
        return self.primitive.num_qubits

    def mul(self, scalar):
        return OperatorStateFn(self.primitive, self.coeff * scalar, self.is_measurement)

    def adjoint(self):
        return OperatorStateFn(self.primitive.adjoint(), np.conj(self.coeff), not self.is_measurement)

    def compose(self, other):
        return ComposedOp([self, other])

    def eval(self, front=None):
        """<front|H|front> (operator_state_fn.py:179-208) through the fused expectation kernel."""
        if front is None:
            return self
        if not self.is_measurement:
            raise ValueError('Cannot compute overlap with StateFn or Operator if not Measurement.')
        # ListOp carve-out of the reference (:199-203): only a plain ListOp (subclasses such as
        # SummedOp / ComposedOp are excluded by its type() test), each element scaled by the list's
        # coefficient BEFORE the evaluation -- so front.coeff enters as |front.coeff|^2
        if type(front) is ListOp:  # pylint: disable=unidiomatic-typecheck
            vecs = [f if isinstance(f, VectorStateFn) else StateFn(f) for f in front.oplist]
            eng = self.primitive._engine_for()
            vals = eng.expect_batch([v.device_vector() for v in vecs])
            return front.combo_fn([complex(np.round(v * self.coeff * abs(s.coeff * front.coeff) ** 2, EVAL_SIG_DIGITS))
                                   for v, s in zip(vals, vecs)])
        if not isinstance(front, VectorStateFn):
            front = StateFn(front)
        val = self.primitive._engine_for().expect(front.device_vector())
        return complex(np.round(val * self.coeff * abs(front.coeff) ** 2, decimals=EVAL_SIG_DIGITS))


class ComposedOp(ListOp):
    """list_ops/composed_op.py:111-128: eval right-to-left."""

    distributive = False

    def compose(self, other):
        return ComposedOp(self.oplist + [other], self.coeff)

    def eval(self, front=None):
        ops = list(self.oplist)
        if front is not None:
            ops.append(front if isinstance(front, OperatorBase) else StateFn(front))
        cur = ops[-1]
        for op in reversed(ops[:-1]):
            cur = op.eval(cur)
        if isinstance(cur, list):
            return [self.coeff * c for c in cur]
        return self.coeff * cur if isinstance(cur, numbers.Number) else cur


def StateFn(primitive, coeff=1.0, is_measurement=False):
    """Factory (state_fns/state_fn.py:47-92): vectors -> VectorStateFn, operators -> OperatorStateFn."""
    if isinstance(primitive, (VectorStateFn, OperatorStateFn)):
        # the reference wraps any OperatorBase in an OperatorStateFn; for an object that already is a
        # state function this shim hands the same state back, with the requested coefficient and
        # measurement flag applied instead of silently dropped
        out = primitive
        if bool(is_measurement) != bool(out.is_measurement):
            out = out.adjoint()
        return out.mul(coeff) if coeff != 1.0 else out
    if hasattr(primitive, 'to_vector_state_fn'):   # CircuitStateFn: prepare psi on the device
        return primitive.to_vector_state_fn()
    if isinstance(primitive, str):                 # basis state '0101' (state_fn.py:66-71)
        primitive = {primitive: 1.0}
    if isinstance(primitive, dict):                # DictStateFn: {bitstring: amplitude}, index = int(b, 2)
        n = len(next(iter(primitive)))             # (state_fns/dict_state_fn.py:148-157)
        vec = np.zeros(1 << n, dtype=np.complex128)
        for bits, amp in primitive.items():
            vec[int(bits, 2)] += amp
        primitive = vec
    if isinstance(primitive, (_PauliSumMixin, WeightedPauliOperator)):
        if isinstance(primitive, WeightedPauliOperator):
            primitive = primitive.to_opflow()
        return OperatorStateFn(primitive, coeff, is_measurement)
    return VectorStateFn(primitive, coeff, is_measurement)


class MatrixExpectation:
    """expectations/matrix_expectation.py:32-49.  ``convert`` keeps the measurement matrix-free
    (the reference converts it to a dense MatrixOp)."""

    def convert(self, operator):
        return operator

    def compute_variance(self, exp_op):
        return 0.0


I = PauliOp(Pauli('I'))  # noqa: E741
X = PauliOp(Pauli('X'))
Y = PauliOp(Pauli('Y'))
Z = PauliOp(Pauli('Z'))
