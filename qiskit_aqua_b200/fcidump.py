"""FCIDump files -> spin-orbital integrals -> (FermionicOperator -> WeightedPauliOperator): the text
format on the input side of the hot path for molecular Hamiltonians.

Drop-in for
  qiskit/chemistry/drivers/fcidumpd/parser.py:24-189      parse()
  qiskit/chemistry/drivers/fcidumpd/dumper.py:20-123      dump()
  qiskit/chemistry/drivers/fcidumpd/fcidumpdriver.py:30-106  FCIDumpDriver
  qiskit/chemistry/qmolecule.py:114-121, 431-534          one/two_body_integrals, onee/twoe_to_spin

The reference walks Python sets of index tuples element by element; here the integral table is read
into arrays once, classified by spin block with index arithmetic, and the permutational symmetry
((ia|jb) = (ai|jb) = (ia|bj) = (jb|ia) ...) is applied as whole-array scatters.
"""
import itertools
import re

import numpy as np

from .operators import AquaError


class QiskitChemistryError(AquaError):
    """qiskit/chemistry/qiskit_chemistry_error.py"""


_INT_KEYS = (('MS2', 0), ('ISYM', 1), ('IPRTIM', -1), ('INT', 5), ('MEMORY', 10000), ('MAXIT', 25), ('NROOT', 1))
_FLOAT_KEYS = (('CORE', 0.0), ('THR', 1e-5), ('THRRES', 0.1))
_NUMBER = r'.*?([-+]?\d*\.\d+|[-+]?\d+),'

# index permutations that leave a same-spin (ia|jb) unchanged, and the subset valid across spins
_PERMS_SAME = ((0, 1, 2, 3), (1, 0, 2, 3), (0, 1, 3, 2), (1, 0, 3, 2),
               (2, 3, 0, 1), (3, 2, 0, 1), (2, 3, 1, 0), (3, 2, 1, 0))
_PERMS_MIXED = _PERMS_SAME[:4]


def _symmetrise(values, given, perms):
    """Fill the entries that the file did not list from a listed permutational partner."""
    out = values.copy()
    filled = given.copy()
    idx = np.argwhere(given)
    vals = values[given]
    for perm in perms[1:]:
        tgt = tuple(idx[:, p] for p in perm)
        free = ~filled[tgt]
        sel = tuple(t[free] for t in tgt)
        out[sel] = vals[free]
        filled[sel] = True
    return out


def parse(fcidump):
    """Parse an FCIDump file into the dictionary the reference's parser returns (same keys:
    NORB, NELEC, MS2, ISYM, ORBSYM, IPRTIM, INT, MEMORY, CORE, MAXIT, THR, THRRES, NROOT, ecore,
    hij, hij_b, hijkl, hijkl_ba, hijkl_bb)."""
    try:
        with open(fcidump, 'r', encoding='utf8') as f:
            text = f.read()
    except OSError as ex:
        raise QiskitChemistryError("Input file '{}' cannot be read!".format(fcidump)) from ex
    end = re.search('(/|&END)', text)
    if end is None:
        raise QiskitChemistryError('The FCIDump namelist is not terminated by / or &END')
    meta = ' '.join(text[:end.start(0)].split())
    out = {}
    for key in ('NORB', 'NELEC'):
        m = re.search(key + _NUMBER, meta)
        if m is None:
            raise QiskitChemistryError('The required {} entry of the FCIDump format is missing!'.format(key))
        out[key] = int(m.group(1))
    norb = out['NORB']
    for key, default in _INT_KEYS:
        m = re.search(key + _NUMBER, meta)
        out[key] = int(m.group(1)) if m else default
    for key, default in _FLOAT_KEYS:
        m = re.search(key + _NUMBER, meta)
        out[key] = float(m.group(1)) if m else default
    m = re.search(r'ORBSYM.*?' + r'(\d+),' * norb, meta)
    out['ORBSYM'] = [int(s) for s in m.groups()] if m else [1] * norb

    rows = [line.split() for line in text[end.end(0):].split('\n') if line.strip()]
    if any(len(r) != 5 for r in rows):
        raise QiskitChemistryError('Every integral line of an FCIDump file has a value and four indices')
    x = np.array([float(r[0]) for r in rows])
    ind = np.array([[int(v) for v in r[1:]] for r in rows], dtype=np.int64).reshape(-1, 4)
    i, a, j, b = ind.T
    # an unrestricted file labels spin orbitals: it has the diagonal 1e entry (2 NORB, 2 NORB)
    uhf = bool(np.any((i == 2 * norb) & (a == 2 * norb) & (j == 0) & (b == 0)))
    top = 2 * norb if uhf else norb

    core = (i == 0) & (a == 0) & (j == 0) & (b == 0)
    if core.any():
        out['ecore'] = float(x[core][-1])
    one = (j == 0) & (b == 0) & (a != 0) & ~core          # (i, 0, 0, 0) lines are MO energies: ignored
    two = ~((j == 0) & (b == 0)) & ~((a == 0) & (j == 0) & (b == 0))
    bad1 = one & ((i < 1) | (a < 1) | (i > top) | (a > top) | ((i > norb) != (a > norb)))
    if bad1.any():
        k = int(np.flatnonzero(bad1)[0])
        raise QiskitChemistryError("Unkown 1-electron integral indices encountered in '{}'".format((int(i[k]), int(a[k]))))
    bad2 = two & ((ind.min(axis=1) < 1) | (ind.max(axis=1) > top) | ((i > norb) != (a > norb)) | ((j > norb) != (b > norb)))
    if bad2.any():
        k = int(np.flatnonzero(bad2)[0])
        raise QiskitChemistryError("Unkown 2-electron integral indices encountered in '{}'".format(tuple(int(v) for v in ind[k])))

    def block1(sel, shift):
        h = np.zeros((norb, norb))
        given = np.zeros((norb, norb), dtype=bool)
        h[i[sel] - 1 - shift, a[sel] - 1 - shift] = x[sel]
        given[i[sel] - 1 - shift, a[sel] - 1 - shift] = True
        return np.where(given, h, h.T)                     # unlisted (p, q) takes the listed (q, p)

    def block2(sel, shift_bra, shift_ket, perms):
        h = np.zeros((norb,) * 4)
        given = np.zeros((norb,) * 4, dtype=bool)
        pos = (i[sel] - 1 - shift_bra, a[sel] - 1 - shift_bra, j[sel] - 1 - shift_ket, b[sel] - 1 - shift_ket)
        h[pos] = x[sel]
        given[pos] = True
        return _symmetrise(h, given, perms)

    bra_beta, ket_beta = i > norb, j > norb
    out['hij'] = block1(one & ~bra_beta, 0)
    out['hijkl'] = block2(two & ~bra_beta & ~ket_beta, 0, 0, _PERMS_SAME)
    out['hij_b'] = out['hijkl_ba'] = out['hijkl_bb'] = None
    if uhf:
        out['hij_b'] = block1(one & bra_beta, norb)
        out['hijkl_bb'] = block2(two & bra_beta & ket_beta, norb, norb, _PERMS_SAME)
        ab = block2(two & ~bra_beta & ket_beta, 0, norb, _PERMS_MIXED)
        ba = block2(two & bra_beta & ~ket_beta, norb, 0, _PERMS_MIXED)
        if np.allclose(ab, 0.0) == np.allclose(ba, 0.0):
            raise QiskitChemistryError('Encountered mixed sets of indices for the 2-electron integrals. '
                                       'Either alpha/beta or beta/alpha matrix should be specified.')
        out['hijkl_ba'] = ab.transpose() if np.allclose(ba, 0.0) else ba
    return out


def _line(value, indices):
    return '{:23.16E}{:4d}{:4d}{:4d}{:4d}\n'.format(value, *indices)


def _unique_2e_lines(h, norb, offsets, mixed):
    """One line per distinct integral: an element is skipped when a permutational partner with a
    (numerically) equal value has been written already (dumper.py:87-119)."""
    perms = _PERMS_MIXED if mixed else _PERMS_SAME
    written = {}
    lines = []
    for elem in (tuple(int(v) for v in e) for e in np.argwhere(~np.isclose(h, 0.0, atol=1e-14))):
        value = h[elem]
        if len(set(elem)) > 1:
            orbit = min(tuple(elem[p] for p in perm) for perm in perms)
            seen = written.setdefault(orbit, [])
            if any(np.isclose(value, v) for v in seen):
                continue
            seen.append(value)
        lines.append(_line(value, (elem[0] + offsets[0], elem[1] + offsets[0], elem[2] + offsets[1], elem[3] + offsets[1])))
    return lines


def _unique_1e_lines(h, norb, offset):
    lines, written = [], set()
    for p, q in itertools.product(range(norb), repeat=2):
        if p != q:
            if (q, p) in written and np.isclose(h[p][q], h[q][p]):
                continue
            written.add((p, q))
        lines.append(_line(h[p][q], (p + offset, q + offset, 0, 0)))
    return lines


def dump(outpath, norb, nelec, hijs, hijkls, einact, ms2=0, orbsym=None, isym=1):
    """Write an FCIDump file (restricted, or unrestricted when the beta blocks are given)."""
    hij, hij_b = hijs
    hijkl, hijkl_ba, hijkl_bb = hijkls
    betas = [hij_b, hijkl_ba, hijkl_bb]
    assert all(h is None for h in betas) or all(h is not None for h in betas)
    assert norb == hij.shape[0] == hijkl.shape[0]
    parts = ['&FCI NORB={:4d},NELEC={:4d},MS2={:4d}\n'.format(norb, nelec, ms2)]
    if orbsym is None:
        parts.append(' ORBSYM=' + '1,' * norb + '\n')
    else:
        assert len(orbsym) == norb
        parts.append(' ORBSYM=' + ','.join(orbsym) + '\n')
    parts.append(' ISYM={:d},\n&END\n'.format(isym))
    parts += _unique_2e_lines(np.asarray(hijkl), norb, (1, 1), False)
    if hijkl_ba is not None:
        parts += _unique_2e_lines(np.asarray(hijkl_ba).transpose(), norb, (1, 1 + norb), True)
    if hijkl_bb is not None:
        parts += _unique_2e_lines(np.asarray(hijkl_bb), norb, (1 + norb, 1 + norb), False)
    parts += _unique_1e_lines(hij, norb, 1)
    if hij_b is not None:
        parts += _unique_1e_lines(hij_b, norb, 1 + norb)
    parts.append(_line(einact, (0, 0, 0, 0)))
    with open(outpath, 'w', encoding='utf8') as f:
        f.write(''.join(parts))


class QMolecule:
    """The fields of the reference's QMolecule that the FCIDump driver fills, plus the conversion
    to spin-orbital integrals (orbital p spin-up -> mode p, spin-down -> mode norb + p)."""

    def __init__(self):
        self.nuclear_repulsion_energy = None
        self.num_orbitals = None
        self.multiplicity = None
        self.molecular_charge = 0
        self.num_alpha = self.num_beta = None
        self.num_atoms = None
        self.atom_symbol = None
        self.atom_xyz = None
        self.mo_onee_ints = self.mo_onee_ints_b = None
        self.mo_eri_ints = self.mo_eri_ints_bb = self.mo_eri_ints_ba = None

    @property
    def one_body_integrals(self):
        return QMolecule.onee_to_spin(self.mo_onee_ints, self.mo_onee_ints_b)

    @property
    def two_body_integrals(self):
        return QMolecule.twoe_to_spin(self.mo_eri_ints, self.mo_eri_ints_bb, self.mo_eri_ints_ba)

    @staticmethod
    def onee_to_spin(mohij, mohij_b=None, threshold=1E-12):
        """h1 in spin orbitals: blockdiag(alpha, beta), entries below the threshold dropped."""
        mohij = np.asarray(mohij)
        mohij_b = mohij if mohij_b is None else np.asarray(mohij_b)
        n = mohij.shape[0]
        h1 = np.zeros((2 * n, 2 * n))
        h1[:n, :n] = np.where(np.abs(mohij) > threshold, mohij, 0.0)
        h1[n:, n:] = np.where(np.abs(mohij_b) > threshold, mohij_b, 0.0)
        return h1

    @staticmethod
    def twoe_to_spin(mohijkl, mohijkl_bb=None, mohijkl_ba=None, threshold=1E-12):
        """h2[p,q,r,s] = -1/2 (chemist -> physicist reordering 'ijkl->ljik') for spin(p) = spin(s),
        spin(q) = spin(r); the four spin blocks are written as slabs."""
        aa = np.einsum('ijkl->ljik', np.asarray(mohijkl))
        if mohijkl_bb is None or mohijkl_ba is None:
            bb = ba = ab = aa
        else:
            bb = np.einsum('ijkl->ljik', np.asarray(mohijkl_bb))
            ba = np.einsum('ijkl->ljik', np.asarray(mohijkl_ba))
            ab = np.einsum('ijkl->ljik', np.asarray(mohijkl_ba).transpose())
        n = aa.shape[0]
        h2 = np.zeros((2 * n,) * 4)
        lo, hi = slice(0, n), slice(n, 2 * n)
        for outer, inner, ints in ((lo, lo, aa), (lo, hi, ba), (hi, lo, ab), (hi, hi, bb)):
            h2[outer, inner, inner, outer] = np.where(np.abs(ints) > threshold, -0.5 * ints, 0.0)
        return h2


class FCIDumpDriver:
    """FCIDumpDriver(path).run() -> QMolecule (fcidumpdriver.py:30-88)."""

    def __init__(self, fcidump_input, atoms=None):
        if not isinstance(fcidump_input, str):
            raise QiskitChemistryError("The fcidump_input must be str, not '{}'".format(fcidump_input))
        self._fcidump_input = fcidump_input
        self.atoms = atoms

    def run(self):
        data = parse(self._fcidump_input)
        mol = QMolecule()
        mol.nuclear_repulsion_energy = data.get('ecore', None)
        mol.num_orbitals = data.get('NORB')
        mol.multiplicity = data.get('MS2', 0) + 1
        mol.num_beta = (data.get('NELEC') - (mol.multiplicity - 1)) // 2
        mol.num_alpha = data.get('NELEC') - mol.num_beta
        if self.atoms is not None:
            mol.num_atoms = len(self.atoms)
            mol.atom_symbol = self.atoms
            mol.atom_xyz = [[float('NaN')] * 3] * len(self.atoms)
        mol.mo_onee_ints = data.get('hij', None)
        mol.mo_onee_ints_b = data.get('hij_b', None)
        mol.mo_eri_ints = data.get('hijkl', None)
        mol.mo_eri_ints_bb = data.get('hijkl_bb', None)
        mol.mo_eri_ints_ba = data.get('hijkl_ba', None)
        return mol

    @staticmethod
    def dump(q_mol, outpath, orbsym=None, isym=1):
        dump(outpath, q_mol.num_orbitals, q_mol.num_alpha + q_mol.num_beta,
             (q_mol.mo_onee_ints, q_mol.mo_onee_ints_b),
             (q_mol.mo_eri_ints, q_mol.mo_eri_ints_ba, q_mol.mo_eri_ints_bb),
             q_mol.nuclear_repulsion_energy, ms2=q_mol.multiplicity - 1, orbsym=orbsym, isym=isym)
