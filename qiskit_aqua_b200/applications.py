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
    """Second-quantised operator -> qubit operator (jordan_wigner, parity, bravyi_kitaev).
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

    @h1.setter
    def h1(self, new_h1):
        self._h1 = new_h1

    @h2.setter
    def h2(self, new_h2):
        self._h2 = new_h2

    def __eq__(self, other):
        return bool(np.all(self._h1 == other._h1) and np.all(self._h2 == other._h2))

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    # -- basis changes and mode reduction (fermionic_operator.py:126-160, 533-617) -----------------
    def transform(self, unitary_matrix):
        """h1 -> U^dag h1 U, h2[i,j,k,l] -> sum conj(U_ia) U_jb conj(U_kc) U_ld h2[i,j,k,l]."""
        u = np.asarray(unitary_matrix)
        self._h1 = u.T.conj().dot(self._h1.dot(u))
        self._h2 = np.einsum('ia,jb,kc,ld,ijkl->abcd', u.conj(), u, u.conj(), u, self._h2, optimize=True)

    def fermion_mode_elimination(self, fermion_mode_array):
        """Drop the listed modes (rows/columns of h1 and h2)."""
        keep = np.setdiff1d(np.arange(self._modes), np.asarray(fermion_mode_array))
        h1_new = self._h1[np.ix_(keep, keep)].copy()
        if np.count_nonzero(self._h2) > 0:
            h2_new = self._h2[np.ix_(keep, keep, keep, keep)].copy()
        else:
            h2_new = np.zeros((len(keep),) * 4)
        return FermionicOperator(h1_new, h2_new)

    def fermion_mode_freezing(self, fermion_mode_array):
        """Treat the listed modes as occupied: their interaction with the kept modes is folded
        into h1, their own energy into the returned shift.  With F the frozen and K the kept modes
        (h2 indexed [i, j, l, k] as in the reference):
            h1[l, j] -= sum_{i in F} h2[i, j, l, i]      h1[l, k] += sum_{i in F} h2[i, i, l, k]
            h1[i, j] += sum_{l in F} h2[i, j, l, l]      h1[i, k] -= sum_{l in F} h2[i, l, l, k]
            shift = sum_{i != l in F} (h2[i, i, l, l] - h2[i, l, l, i]) + sum_{i in F} h1[i, i]
        """
        frozen = np.sort(np.asarray(fermion_mode_array))
        keep = np.setdiff1d(np.arange(self._modes), frozen)
        h_1 = self._h1.copy()
        h2 = self._h2
        energy_shift = 0.0
        h2_new = np.zeros((len(keep),) * 4)
        if np.count_nonzero(h2) > 0:
            h2_new = h2[np.ix_(keep, keep, keep, keep)].copy()
            kk = np.ix_(keep, keep)
            f = frozen
            # h2[i, j, l, i] summed over frozen i -> [j, l]; transposed onto h1[l, j]
            h_1[kk] -= np.einsum('ijli->lj', h2[np.ix_(f, keep, keep, f)])
            h_1[kk] += np.einsum('iilk->lk', h2[np.ix_(f, f, keep, keep)])
            h_1[kk] += np.einsum('ijll->ij', h2[np.ix_(keep, keep, f, f)])
            h_1[kk] -= np.einsum('illk->ik', h2[np.ix_(keep, f, f, keep)])
            ff = h2[np.ix_(f, f, f, f)]
            coulomb = np.einsum('iill->il', ff)
            exchange = np.einsum('illi->il', ff)
            off = ~np.eye(len(f), dtype=bool)
            energy_shift += float(np.sum(coulomb[off]) - np.sum(exchange[off]))
        energy_shift += np.sum(np.diagonal(h_1)[frozen])
        return FermionicOperator(h_1[np.ix_(keep, keep)], h2_new), energy_shift

    # -- observables used as aux operators of the eigensolver (fermionic_operator.py:619-760) -------
    def total_particle_number(self):
        n = self._modes
        return FermionicOperator(np.eye(n, dtype=complex), np.zeros((n,) * 4))

    def total_magnetization(self):
        """(N_up - N_down)/2 in block spin order (first half up)."""
        n = self._modes
        h1 = np.eye(n, dtype=complex) * 0.5
        h1[n // 2:, n // 2:] *= -1.0
        return FermionicOperator(h1, np.zeros((n,) * 4))

    def total_angular_momentum(self):
        """S^2 = S_x^2 + S_y^2 + S_z^2 for block spin order, assembled with index arrays:
        orbital p has the modes (p, p + m), m = modes / 2."""
        n = self._modes
        m = n // 2
        h1 = np.zeros((n, n))
        h2 = np.zeros((n,) * 4)
        p, q = np.indices((m, m)).reshape(2, -1)
        d = p == q
        a, b = p[~d], q[~d]                      # different orbitals
        o = p[d]                                 # same orbital
        # S_x^2 + S_y^2: the spin-flip pairs; the p(up)p(down) q(up)q(down) terms cancel between x and y
        h2[a + m, a, b, b + m] += 2.0
        h2[a, a + m, b + m, b] += 2.0
        # S_z^2
        h2[a, a, b, b] += 1.0
        h2[a + m, a + m, b, b] -= 1.0
        h2[a, a, b + m, b + m] -= 1.0
        h2[a + m, a + m, b + m, b + m] += 1.0
        # same orbital: number-operator products
        h2[o, o, o + m, o + m] -= 2.0
        h2[o + m, o + m, o, o] -= 2.0
        h2[o, o + m, o + m, o] += 1.0
        h2[o + m, o, o, o + m] += 1.0
        h1[o, o] += 3.0
        h1[o + m, o + m] += 3.0
        return FermionicOperator(h1=h1 * 0.25, h2=h2 * 0.25)

    @staticmethod
    def _mul(a, b):
        """(x,z,phase) product of two label Paulis given as (x, z)."""
        cx, cz = a[0] ^ b[0], a[1] ^ b[1]
        e = (bin(a[0] & a[1]).count('1') + bin(b[0] & b[1]).count('1') - bin(cx & cz).count('1')
             + 2 * bin(a[1] & b[0]).count('1')) % 4
        return (cx, cz), (1, 1j, -1, -1j)[e]

    @staticmethod
    def _mode_paulis(map_type, n):
        """Per mode i the two Paulis (A_i, B_i) as (x_mask, z_mask), adag_i = (A_i - i B_i)/2,
        a_i = (A_i + i B_i)/2 (fermionic_operator.py:163-342).

          jordan_wigner  A = Z_0..Z_{i-1} X_i                    B = Z_0..Z_{i-1} Y_i
          parity         A = Z_{i-1} X_i X_{i+1}..X_{n-1}        B = Y_i X_{i+1}..X_{n-1}
          bravyi_kitaev  A = X_{U(i)} X_i Z_{P(i)}               B = X_{U(i)} Y_i Z_{P(i) minus F(i)}
        with the update / parity / flip sets of the Fenwick tree over the next power of two,
        restricted to the n modes present.
        """
        if map_type == 'jordan_wigner':
            return [((1 << i, (1 << i) - 1), (1 << i, (1 << (i + 1)) - 1)) for i in range(n)]
        if map_type == 'parity':
            upper = (1 << n) - 1
            return [((upper & ~((1 << i) - 1), (1 << (i - 1)) if i else 0),
                     (upper & ~((1 << i) - 1), 1 << i)) for i in range(n)]
        if map_type == 'bravyi_kitaev':
            size = 1
            while size < n:
                size *= 2
            keep = (1 << n) - 1
            modes = []
            for j in range(n):
                update = parity = flip = 0
                lo, span = 0, size                  # walk down the halving tree towards leaf j
                while span > 1:
                    half = span // 2
                    if j - lo < half:
                        update |= 1 << (lo + span - 1)      # the block's top node also stores this half
                    else:
                        parity |= 1 << (lo + half - 1)      # everything left of j inside this block
                        if j == lo + span - 1:              # ... and j itself stores that left half
                            flip |= 1 << (lo + half - 1)
                        lo += half
                    span = half
                update &= keep
                parity &= keep
                flip &= keep
                modes.append(((update | (1 << j), parity), (update | (1 << j), (parity & ~flip) | (1 << j))))
            return modes
        raise ValueError('Please specify the supported modes: jordan_wigner, parity, bravyi_kitaev, '
                         'bksf')

    def mapping(self, map_type, threshold=1e-8):
        map_type = map_type.lower()
        if map_type == 'bksf':
            raise ValueError('the bksf (graph-based) encoding is not provided on this path')
        n = self._modes
        modes = self._mode_paulis(map_type, n)
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
