"""Input generators either side of the hot path, restated on bit masks (no Python 2^n loops).

  * FermionicOperator(h1, h2).mapping('jordan_wigner')
      qiskit/chemistry/fermionic_operator.py:163-190, 344-474
  * the Ising-problem generators (max_cut, ..., random_graph, sample_most_likely) live in
    ising.py and are re-exported here
"""
import itertools

import numpy as np

from .ising import max_cut, random_graph, sample_most_likely  # noqa: F401
from .operators import Pauli, WeightedPauliOperator


class FermionicOperator:
    """Second-quantised operator -> qubit operator (Jordan-Wigner only on this path).
    h2 uses the class's own convention h2[i,j,k,m] -> adag_i adag_k a_m a_j."""

    def __init__(self, h1, h2=None, ph_trans_shift=None):
        self._h1 = np.asarray(h1)
        n = self._h1.shape[0]
        self._h2 = np.zeros((n, n, n, n), dtype=self._h1.dtype) if h2 is None else np.asarray(h2)
        self._ph_trans_shift = ph_trans_shift
        self._modes = n

    @property
    def modes(self):
        return self._modes

    @property
    def h1(self):
        return self._h1

    @property
    def h2(self):
        return self._h2

    @staticmethod
    def _mul(a, b):
        """(x,z,phase) product of two label Paulis given as (x, z)."""
        cx, cz = a[0] ^ b[0], a[1] ^ b[1]
        e = (bin(a[0] & a[1]).count('1') + bin(b[0] & b[1]).count('1') - bin(cx & cz).count('1')
             + 2 * bin(a[1] & b[0]).count('1')) % 4
        return (cx, cz), (1, 1j, -1, -1j)[e]

    def mapping(self, map_type, threshold=1e-8):
        if map_type.lower() != 'jordan_wigner':
            raise ValueError('only the jordan_wigner mapping is provided on this path')
        n = self._modes
        # adag_i = (A_i - i B_i)/2, a_i = (A_i + i B_i)/2 with A = Z..ZX, B = Z..ZY
        modes = [((1 << i, (1 << i) - 1), (1 << i, (1 << (i + 1)) - 1)) for i in range(n)]
        acc, order = {}, []

        def add(key, w):
            if key not in acc:
                acc[key] = 0.0
                order.append(key)
            acc[key] += w

        def chop():
            for key in list(order):
                w = complex(acc[key])
                re = w.real if abs(w.real) >= threshold else 0.0
                im = w.imag if abs(w.imag) >= threshold else 0.0
                if re == 0.0 and im == 0.0:
                    del acc[key]
                    order.remove(key)
                else:
                    acc[key] = re + 1j * im

        def emit(local, lorder):
            for key in lorder:
                if local[key] != 0:
                    add(key, local[key])

        for i, j in zip(*np.nonzero(self._h1)):
            v = self._h1[i, j]
            local, lorder = {}, []
            for al, be in itertools.product(range(2), repeat=2):
                p, ph = self._mul(modes[i][al], modes[j][be])
                c = v / 4 * ph * (-1j) ** al * (1j) ** be
                if abs(c) > threshold:
                    if p not in local:
                        local[p] = 0.0
                        lorder.append(p)
                    local[p] += c
            emit(local, lorder)
        chop()
        for i, j, k, m in zip(*np.nonzero(self._h2)):
            v = self._h2[i, j, k, m]
            local, lorder = {}, []
            for al, be, ga, de in itertools.product(range(2), repeat=4):
                p1, f1 = self._mul(modes[i][al], modes[k][be])
                p2, f2 = self._mul(p1, modes[m][ga])
                p3, f3 = self._mul(p2, modes[j][de])
                c = v / 16 * (f1 * f2 * f3) * (-1j) ** (al + be) * (1j) ** (ga + de)
                if abs(c) > threshold:
                    if p3 not in local:
                        local[p3] = 0.0
                        lorder.append(p3)
                    local[p3] += c
            emit(local, lorder)
        chop()
        paulis = [[acc[key], Pauli.from_masks(key[0], key[1], n)] for key in order]
        op = WeightedPauliOperator(paulis=paulis)
        if self._ph_trans_shift is not None:
            op = op + WeightedPauliOperator(paulis=[[self._ph_trans_shift, Pauli('I' * n)]])
        return op
