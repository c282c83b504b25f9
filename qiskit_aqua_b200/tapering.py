"""Z2-symmetry tapering of a Pauli-sum Hamiltonian: fewer qubits for the same spectrum sector, i.e. a
2^k times smaller statevector for the hot path.  Drop-in for `Z2Symmetries`
(qiskit/aqua/operators/legacy/weighted_pauli_operator.py:960-1330) and `kernel_F2`
(legacy/common.py:117-175), on bit masks: the GF(2) elimination works on Python-int bit rows, the
single-qubit partners are found with mask tests, and tapering deletes bits with two shifts.
"""
import copy
import itertools

import numpy as np

from .operators import AquaError, Pauli, WeightedPauliOperator


def _drop_bits(mask, positions):
    """Remove the given bit positions from a mask (higher bits move down)."""
    for q in sorted(set(positions), reverse=True):      # np.delete semantics: duplicates count once
        mask = ((mask >> (q + 1)) << q) | (mask & ((1 << q) - 1))
    return mask


def kernel_F2(rows, width):  # noqa: N802
    """Null space over GF(2) of the matrix whose rows are the bit masks `rows` (bit j = column j),
    with the reference's elimination order, so the SAME basis comes out: Gauss-Jordan on the
    columns taken in order, pivot = first row index with a one (legacy/common.py:117-175).
    Returns masks over the `width` columns."""
    m = len(rows)
    # augmented "column vectors": low m bits = the column of the matrix, then the identity part
    cols = []
    for j in range(width):
        v = 0
        for i, r in enumerate(rows):
            v |= ((r >> j) & 1) << i
        cols.append(v | (1 << (m + j)))
    for i in range(width):
        pivot = (cols[i] & -cols[i]).bit_length() - 1 if cols[i] else 0
        for k in range(width):
            if k != i and (cols[k] >> pivot) & 1:
                cols[k] ^= cols[i]
    low = (1 << m) - 1
    return [c >> m for c in cols if (c & low) == 0 and (c >> m) != 0]


class Z2Symmetries:
    """Symmetries S_k (Paulis commuting with every term), single-qubit partners sigma_k acting on
    qubit sq_list[k], and the eigenvalue sector (+1/-1 per symmetry)."""

    def __init__(self, symmetries, sq_paulis, sq_list, tapering_values=None):
        if len(symmetries) != len(sq_paulis):
            raise AquaError("Number of Z2 symmetries has to be the same as number of single-qubit pauli x.")
        if len(sq_paulis) != len(sq_list):
            raise AquaError("Number of single-qubit pauli x has to be the same as length of single-qubit list.")
        if tapering_values is not None and len(sq_list) != len(tapering_values):
            raise AquaError("The length of single-qubit list has to be the same as length of tapering values.")
        self._symmetries = symmetries
        self._sq_paulis = sq_paulis
        self._sq_list = sq_list
        self._tapering_values = tapering_values

    symmetries = property(lambda self: self._symmetries)
    sq_paulis = property(lambda self: self._sq_paulis)
    sq_list = property(lambda self: self._sq_list)

    @property
    def tapering_values(self):
        return self._tapering_values

    @tapering_values.setter
    def tapering_values(self, new_value):
        self._tapering_values = new_value

    @property
    def cliffords(self):
        """U_k = (S_k + sigma_k)/sqrt 2; U_k S_k U_k = sigma_k."""
        return [WeightedPauliOperator(paulis=[[1 / np.sqrt(2), s], [1 / np.sqrt(2), x]])
                for s, x in zip(self._symmetries, self._sq_paulis)]

    def copy(self):
        return copy.deepcopy(self)

    def is_empty(self):
        return not (self._symmetries != [] and self._sq_paulis != [] and self._sq_list != [])

    def __str__(self):
        ret = ['Z2 symmetries:', 'Symmetries:'] + [s.to_label() for s in self._symmetries]
        ret += ['Single-Qubit Pauli X:'] + [x.to_label() for x in self._sq_paulis]
        ret += ['Cliffords:'] + [c.print_details() for c in self.cliffords]
        ret += ['Qubit index:', str(self._sq_list), 'Tapering values:']
        if self._tapering_values is None:
            values = ', '.join(str(list(c)) for c in itertools.product([1, -1], repeat=len(self._sq_list)))
            ret.append('  - Possible values: ' + values)
        else:
            ret.append(str(self._tapering_values))
        return '\n'.join(ret)

    # -- discovery ------------------------------------------------------------------------------
    @classmethod
    def find_Z2_symmetries(cls, operator):  # noqa: N802
        """All independent Paulis commuting with every term (kernel of the symplectic check
        matrix), each with a single-qubit Pauli that anticommutes with it and commutes with the
        other symmetries."""
        if operator.is_empty():
            return cls([], [], [], None)
        n = operator.num_qubits
        # row of a term: x bits then z bits; a kernel vector pairs its low half with x, i.e. the low
        # half is the symmetry's z mask and the high half its x mask
        rows = [p.x_mask | (p.z_mask << n) for _, p in operator.paulis]
        kernel = kernel_F2(rows, 2 * n)
        if not kernel:
            return cls([], [], [], None)
        low = (1 << n) - 1
        sym = [(v & low, v >> n) for v in kernel]                 # (z mask, x mask)
        symmetries, sq_paulis, sq_list = [], [], []
        for r, (z, x) in enumerate(sym):
            symmetries.append(Pauli.from_masks(x, z, n))
            others_z = others_x = others_diff = 0
            for k, (oz, ox) in enumerate(sym):
                if k != r:
                    others_z |= oz
                    others_x |= ox
                    others_diff |= oz ^ ox
            for q in range(n):
                bit = 1 << q
                if not others_z & bit and z & bit:                # X_q: only this symmetry has Z/Y here
                    sq_paulis.append(Pauli.from_masks(bit, 0, n))
                elif not others_x & bit and x & bit:              # Z_q
                    sq_paulis.append(Pauli.from_masks(0, bit, n))
                elif not others_diff & bit and (z ^ x) & bit:     # Y_q
                    sq_paulis.append(Pauli.from_masks(bit, bit, n))
                else:
                    continue
                sq_list.append(q)
                break
            else:
                raise AquaError('no single-qubit Pauli separates symmetry %d from the others' % r)
        return cls(symmetries, sq_paulis, sq_list, None)

    # -- tapering -------------------------------------------------------------------------------
    def taper(self, operator, tapering_values=None):
        """Rotate with the Cliffords, replace sigma_k by its eigenvalue and delete qubit sq_list[k].
        With no sector given (here or at construction) all 2^k sectors are returned."""
        if not self._symmetries or not self._sq_paulis or not self._sq_list:
            raise AquaError("Z2 symmetries, single qubit pauli and single qubit list cannot be empty.")
        if operator.is_empty():
            return operator
        for clifford in self.cliffords:
            operator = clifford * operator * clifford
        tapering_values = tapering_values if tapering_values is not None else self._tapering_values
        n = operator.num_qubits
        touched = [1 << q for q in self._sq_list]
        terms = [(w, p.x_mask, p.z_mask) for w, p in operator.paulis]

        def one_sector(values):
            z2 = self.copy()
            z2.tapering_values = values
            paulis = []
            for w, x, z in terms:
                if np.real(w) == 0.0 and np.imag(w) == 0.0:
                    continue                      # cancelled by the Clifford rotation: contributes no term
                for bit, value in zip(touched, values):
                    if (x | z) & bit:
                        w = value * w
                paulis.append([w, Pauli.from_masks(_drop_bits(x, self._sq_list), _drop_bits(z, self._sq_list),
                                                   n - len(set(self._sq_list)))])
            return WeightedPauliOperator(paulis=paulis, z2_symmetries=z2, name=operator.name)

        if tapering_values is None:
            return [one_sector(list(c)) for c in itertools.product([1, -1], repeat=len(self._sq_list))]
        return one_sector(tapering_values)

    @staticmethod
    def two_qubit_reduction(operator, num_particles):
        """Parity / binary-tree mapped fermionic Hamiltonians in block spin order: the middle and
        the last qubit hold the alpha and the total particle-number parity; taper both."""
        if operator.is_empty():
            return operator
        if isinstance(num_particles, (tuple, list)):
            num_alpha, num_beta = num_particles[0], num_particles[1]
        else:
            num_alpha = num_beta = num_particles // 2
        values = [1 if num_alpha % 2 == 0 else -1, 1 if (num_alpha + num_beta) % 2 == 0 else -1]
        n = operator.num_qubits
        sq_list = [n // 2 - 1, n - 1]
        symmetries = [Pauli.from_masks(0, 1 << q, n) for q in sq_list]
        sq_paulis = [Pauli.from_masks(1 << q, 0, n) for q in sq_list]
        return Z2Symmetries(symmetries, sq_paulis, sq_list, values).taper(operator)

    def consistent_tapering(self, operator):
        """Taper another operator (e.g. an aux operator) the way this sector was tapered."""
        if operator.is_empty():
            raise AquaError("Can not taper an empty operator.")
        for s in self._symmetries:
            for _, p in operator.paulis:
                if (bin(p.x_mask & s.z_mask).count('1') + bin(p.z_mask & s.x_mask).count('1')) % 2:
                    raise AquaError("The given operator does not commute with the symmetry, can not taper it.")
        return self.taper(operator)
