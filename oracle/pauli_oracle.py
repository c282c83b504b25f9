"""CPU oracle for the Pauli-sum H.psi / <psi|H|psi> hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of what qiskit-aqua 0.10.0 computes on the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
(``qiskit_aqua_b200``) never does.

Parity status: PINNED (i) by the reference's own golden vectors (SURVEY.md section 8c, G1..G13;
``tests/test_oracle_golden.py``) and (ii) by outputs of the reference's OWN code run in the build
container -- WeightedPauliOperator / op_converter / MatrixOperator / max_cut / random_graph /
FermionicOperator / NumPyEigensolver / NumPyMinimumEigensolver with only terra's Pauli stubbed (``tests/golden/make_reference_fixtures.py`` ->
``tests/golden/reference_run.json``, checked by ``tests/test_reference_run.py``).  The per-Pauli CSR constructor itself lives in the
un-vendored dependency ``qiskit-terra>=0.18.0`` (reference ``requirements.txt:1``), which is
not installed here, so its published convention is restated:

    P[i, i ^ x] = (-i)^{popc(x & z)} * (-1)^{popc(i & z)}          (row index in the parity)

with qubit q <-> bit q of the basis index, label strings big-endian (leftmost char = qubit n-1)
and ``Pauli((z, x))`` arrays little-endian.  That convention is pinned from inside Aqua by
``test/aqua/operators/test_op_construction.py:95-116,158-162`` (G4, G9).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import itertools
from typing import Iterable, List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

Term = Tuple[complex, str]  # (weight, big-endian label)


# --------------------------------------------------------------------------------------------
# Pauli <-> masks (terra convention; pinned by G4/G9)
# --------------------------------------------------------------------------------------------
def label_to_masks(label: str) -> Tuple[int, int]:
    """Big-endian label -> (x_mask, z_mask) python ints. Leftmost char is qubit n-1."""
    n = len(label)
    x = z = 0
    for pos, ch in enumerate(label):
        q = n - 1 - pos
        if ch == 'X':
            x |= 1 << q
        elif ch == 'Z':
            z |= 1 << q
        elif ch == 'Y':
            x |= 1 << q
            z |= 1 << q
        elif ch != 'I':
            raise ValueError('bad Pauli char %r' % ch)
    return x, z


def masks_to_label(x: int, z: int, n: int) -> str:
    out = []
    for q in range(n - 1, -1, -1):
        xb, zb = (x >> q) & 1, (z >> q) & 1
        out.append('IXZY'[xb + 2 * zb])
    return ''.join(out)


def zx_arrays_to_label(z: Sequence[bool], x: Sequence[bool]) -> str:
    """terra ``Pauli((z, x))``: arrays are little-endian (index 0 = qubit 0);
    cf. reference qiskit/optimization/applications/ising/max_cut.py:48-52."""
    n = len(z)
    xm = sum(1 << q for q in range(n) if x[q])
    zm = sum(1 << q for q in range(n) if z[q])
    return masks_to_label(xm, zm, n)


def _parity(v: np.ndarray) -> np.ndarray:
    """popcount parity of a uint64 array."""
    v = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(s)
    return (v & np.uint64(1)).astype(np.int8)


def pauli_csr(label: str) -> sp.csr_matrix:
    """terra ``Pauli.to_matrix(sparse=True)`` restated: one non-zero per row.
    Call sites in the reference: legacy/op_converter.py:122, primitive_ops/pauli_op.py:153,164."""
    n = len(label)
    x, z = label_to_masks(label)
    dim = 1 << n
    rows = np.arange(dim, dtype=np.uint64)
    cols = rows ^ np.uint64(x)
    ny = bin(x & z).count('1')
    sign = 1.0 - 2.0 * _parity(rows & np.uint64(z))
    data = ((-1j) ** ny) * sign.astype(np.complex128)
    return sp.csr_matrix((data, cols.astype(np.int64), np.arange(dim + 1, dtype=np.int64)),
                         shape=(dim, dim))


# --------------------------------------------------------------------------------------------
# WeightedPauliOperator container semantics
# --------------------------------------------------------------------------------------------
def simplify(terms: Iterable[Term]) -> List[Term]:
    """legacy/weighted_pauli_operator.py:341-400: merge equal labels in first-seen order, then
    chop(0.0) which drops exactly-zero weights (:399, :419-479)."""
    table = {}
    out: List[List] = []
    for w, lab in terms:
        idx = table.get(lab)
        if idx is None:
            table[lab] = len(out)
            out.append([w, lab])
        else:
            out[idx][0] = out[idx][0] + w
    return chop([(w, lab) for w, lab in out], 0.0)


def chop(terms: Iterable[Term], threshold: float) -> List[Term]:
    """legacy/weighted_pauli_operator.py:419-479 (component-wise real/imag threshold)."""
    out = []
    for w, lab in terms:
        w = complex(w)
        re = w.real if abs(w.real) >= threshold else 0.0
        im = w.imag if abs(w.imag) >= threshold else 0.0
        if re == 0.0 and im == 0.0:
            continue
        out.append((re + 1j * im, lab))
    return out


def from_dict(dictionary: dict) -> List[Term]:
    """legacy/weighted_pauli_operator.py:529-582."""
    terms = []
    for op in dictionary['paulis']:
        c = op['coeff']
        w = complex(c['real'], c['imag']) if 'imag' in c else c['real']
        terms.append((w, op['label']))
    return simplify(terms)


# --------------------------------------------------------------------------------------------
# Literal reference arithmetic (CSR sum -> dot -> vdot)
# --------------------------------------------------------------------------------------------
def ref_to_spmatrix(terms: Sequence[Term]) -> sp.csr_matrix:
    """legacy/op_converter.py:120-123 (``hamiltonian += weight * pauli.to_matrix(sparse=True)``)
    == list_ops/list_op.py:328-329 + summed_op.py:46-49 for a SummedOp of PauliOps."""
    ham = 0
    for w, lab in terms:
        ham = ham + w * pauli_csr(lab)
    return sp.csr_matrix(ham)


def ref_evaluate_with_statevector(terms: Sequence[Term], psi: np.ndarray):
    """legacy/weighted_pauli_operator.py:605-623 -> (np.vdot(psi, H.dot(psi)), 0.0)."""
    if not terms:
        raise ValueError('Operator is empty, check the operator.')
    mat = ref_to_spmatrix(terms)
    return np.vdot(psi, mat.dot(psi)), 0.0


def ref_matrix_expectation(terms: Sequence[Term], psi: np.ndarray) -> complex:
    """opflow chain matrix_expectation.py:32-49 -> operator_state_fn.py:179-208 ->
    matrix_op.py:207 (dense H @ psi) -> vector_state_fn.py:193-196 (np.dot with the bra,
    rounded to EVAL_SIG_DIGITS=18, operator_globals.py:33).  Small n only (dense)."""
    dense = ref_to_spmatrix(terms).toarray()
    hpsi = dense @ psi
    return complex(np.round(np.dot(np.conj(psi), hpsi), decimals=18))


def ref_solve(terms: Sequence[Term], k: int = 1, v0: np.ndarray | None = None):
    """algorithms/eigen_solvers/numpy_eigen_solver.py:159-182.
    Returns (eigvals[k], eigvecs[k, N]) exactly like ``_ret['eigvals'], _ret['eigvecs']``."""
    n = len(terms[0][1])
    sp_mat = ref_to_spmatrix(terms)
    k = min(k, 1 << n)
    if sp.csr_matrix(sp_mat.diagonal()).nnz == sp_mat.nnz:          # :163 diagonal shortcut
        diag = sp_mat.diagonal()
        eigval = np.sort(diag)[:k]
        temp = np.argsort(diag)[:k]
        eigvec = np.zeros((sp_mat.shape[0], k))
        for i, idx in enumerate(temp):
            eigvec[idx, i] = 1.0
    elif k >= (1 << n) - 1:                                            # :171-173 dense
        eigval, eigvec = np.linalg.eig(sp_mat.toarray())
    else:                                                              # :175-176 ARPACK
        eigval, eigvec = spla.eigs(sp_mat, k=k, which='SR', v0=v0)
    if k > 1:
        idx = eigval.argsort()
        eigval = eigval[idx]
        eigvec = eigvec[:, idx]
    return eigval, eigvec.T


def ref_eval_aux(aux_terms: Sequence[Term], wavefn: np.ndarray, threshold: float = 1e-12):
    """numpy_eigen_solver.py:208-226: mat.dot(v).dot(conj(v)), real part, |.|<=1e-12 -> 0."""
    mat = ref_to_spmatrix(aux_terms)
    value = mat.dot(wavefn).dot(np.conj(wavefn))
    value = value.real if abs(value.real) > threshold else 0.0
    return (value, 0)


# --------------------------------------------------------------------------------------------
# Matrix-free restatement (same arithmetic, no CSR) -- usable at sizes the CSR cannot reach
# --------------------------------------------------------------------------------------------
def pack_terms(terms: Sequence[Term]):
    """(weight, label) -> x_mask u64[T], z_mask u64[T], ypow u8[T] (= nY mod 4), coeff c128[T]."""
    xs, zs, ys, cs = [], [], [], []
    for w, lab in terms:
        x, z = label_to_masks(lab)
        xs.append(x)
        zs.append(z)
        ys.append(bin(x & z).count('1') & 3)
        cs.append(complex(w))
    return (np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64),
            np.array(ys, dtype=np.uint8), np.array(cs, dtype=np.complex128))


def mf_apply(n: int, x_mask, z_mask, ypow, coeff, psi: np.ndarray,
             rows: np.ndarray | None = None) -> np.ndarray:
    """y[i] = sum_k c_k (-i)^{nY_k} (-1)^{popc(i & z_k)} psi[i ^ x_k]  (SURVEY.md section 8).
    ``rows`` restricts the output to a subset of row indices (for sampled checks at large n)."""
    idx = np.arange(1 << n, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64)
    y = np.zeros(idx.shape[0], dtype=np.complex128)
    phase = np.array([1, -1j, -1, 1j], dtype=np.complex128)
    for x, z, yp, c in zip(x_mask, z_mask, ypow, coeff):
        sgn = 1.0 - 2.0 * _parity(idx & np.uint64(z))
        y += (c * phase[int(yp) & 3]) * sgn * psi[(idx ^ np.uint64(x)).astype(np.int64)]
    return y


def mf_expect(n, x_mask, z_mask, ypow, coeff, psi) -> complex:
    return complex(np.vdot(psi, mf_apply(n, x_mask, z_mask, ypow, coeff, psi)))


def product_state(n: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """|psi> = prod_q (a_q |0> + b_q |1>) as a dense vector (qubit q <-> bit q); small n only."""
    psi = np.ones(1, dtype=np.complex128)
    for q in range(n):
        psi = np.concatenate([a[q] * psi, b[q] * psi])
    return psi


def product_state_expect(n: int, x_mask, z_mask, ypow, coeff, a: np.ndarray, b: np.ndarray) -> complex:
    """<psi|H|psi> for the product state above WITHOUT any 2^n object: a size-independent parity
    property of the path.  A label Pauli factorises, <P> = prod_q <sigma_q>, with
    <X> = 2 Re(conj(a) b), <Y> = 2 Im(conj(a) b), <Z> = |a|^2 - |b|^2; a packed term is
    c (-i)^{ypow - nY} times its label Pauli (same convention as mf_apply)."""
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    ex = 2.0 * (np.conj(a) * b).real
    ey = 2.0 * (np.conj(a) * b).imag
    ez = np.abs(a) ** 2 - np.abs(b) ** 2
    norm = float(np.prod(np.abs(a) ** 2 + np.abs(b) ** 2))
    phase = np.array([1, -1j, -1, 1j], dtype=np.complex128)
    total = 0.0 + 0.0j
    for x, z, yp, c in zip(x_mask, z_mask, ypow, coeff):
        x, z = int(x), int(z)
        v = 1.0
        ny = 0
        for q in range(n):
            xb, zb = (x >> q) & 1, (z >> q) & 1
            if xb and zb:
                v *= ey[q]
                ny += 1
            elif xb:
                v *= ex[q]
            elif zb:
                v *= ez[q]
            else:
                v *= abs(a[q]) ** 2 + abs(b[q]) ** 2
        total += complex(c) * phase[(int(yp) - ny) & 3] * v
    return total if norm else 0.0


def mf_diag(n, z_mask, coeff) -> np.ndarray:
    """diagonal of an all-Z sum (branch numpy_eigen_solver.py:163-169)."""
    idx = np.arange(1 << n, dtype=np.uint64)
    d = np.zeros(1 << n, dtype=np.complex128)
    for z, c in zip(z_mask, coeff):
        d += c * (1.0 - 2.0 * _parity(idx & np.uint64(z)))
    return d


# --------------------------------------------------------------------------------------------
# Input generators on the path's upstream side
# --------------------------------------------------------------------------------------------
def random_graph(n, weight_range=10, edge_prob=0.3, negative_weight=True, seed=None):
    """qiskit/optimization/applications/ising/common.py:23-60 with
    aqua_globals.random == np.random.default_rng(seed) (aqua_globals.py:89-96)."""
    rng = np.random.default_rng(seed)
    w = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() <= edge_prob:
                w[i, j] = rng.integers(1, weight_range)
                if rng.random() >= 0.5 and negative_weight:
                    w[i, j] *= -1
    w += w.T
    return w


def max_cut_operator(weight_matrix: np.ndarray):
    """qiskit/optimization/applications/ising/max_cut.py:31-54 -> (terms, shift)."""
    n = weight_matrix.shape[0]
    terms = []
    shift = 0.0
    for i in range(n):
        for j in range(i):
            if weight_matrix[i, j] != 0:
                z = np.zeros(n, dtype=bool)
                x = np.zeros(n, dtype=bool)
                z[i] = True
                z[j] = True
                terms.append((0.5 * weight_matrix[i, j], zx_arrays_to_label(z, x)))
                shift -= 0.5 * weight_matrix[i, j]
    return simplify(terms), shift


def sample_most_likely(state_vector: np.ndarray) -> np.ndarray:
    """common.py:146-170, ndarray branch."""
    n = int(np.log2(state_vector.shape[0]))
    k = int(np.argmax(np.abs(state_vector)))
    x = np.zeros(n)
    for i in range(n):
        x[i] = k % 2
        k >>= 1
    return x


def pauli_dot(a: Tuple[int, int], b: Tuple[int, int]) -> Tuple[Tuple[int, int], complex]:
    """Product of two label Paulis A.B = phase * C (C a label Pauli).  Pinned by G12
    (test_weighted_pauli_operator.py:98-112: 0.5 IXYZ * 0.5 ZYIX = -0.25 ZZYY)."""
    ax, az = a
    bx, bz = b
    cx, cz = ax ^ bx, az ^ bz
    # W(x,z) = i^{x.z} X^x Z^z ;  X^a Z^b X^c Z^d = (-1)^{b.c} X^{a^c} Z^{b^d}
    e = (bin(ax & az).count('1') + bin(bx & bz).count('1') - bin(cx & cz).count('1')
         + 2 * bin(az & bx).count('1')) % 4
    return (cx, cz), (1j) ** e


def jordan_wigner(h1: np.ndarray, h2: np.ndarray | None = None, threshold: float = 1e-8):
    """qiskit/chemistry/fermionic_operator.py:163-190 (JW modes), :344-415 (mapping),
    :418-474 (one/two-body products).  h2 in the class's own index convention:
    h2[i,j,k,m] -> adag_i adag_k a_m a_j.  Returns simplified (weight, label) terms."""
    n = h1.shape[0]
    a = []
    for i in range(n):
        zlow = (1 << i) - 1
        a.append(((1 << i, zlow), (1 << i, zlow | (1 << i))))     # (Z..Z X, Z..Z Y)
    acc = {}
    order = []

    def add(w, p):
        if p not in acc:
            acc[p] = 0.0
            order.append(p)
        acc[p] += w

    def chop_acc():
        for p in list(order):
            w = complex(acc[p])
            re = w.real if abs(w.real) >= threshold else 0.0
            im = w.imag if abs(w.imag) >= threshold else 0.0
            if re == 0.0 and im == 0.0:
                del acc[p]
                order.remove(p)
            else:
                acc[p] = re + 1j * im

    for i, j in itertools.product(range(n), repeat=2):
        if h1[i, j] == 0:
            continue
        local = {}
        lorder = []
        for alpha in range(2):
            for beta in range(2):
                p, ph = pauli_dot(a[i][alpha], a[j][beta])
                coeff = h1[i, j] / 4 * ph * (-1j) ** alpha * (1j) ** beta
                if abs(coeff) > threshold:
                    if p not in local:
                        local[p] = 0.0
                        lorder.append(p)
                    local[p] += coeff
        for p in lorder:
            if local[p] != 0:
                add(local[p], p)
    chop_acc()
    if h2 is not None:
        for i, j, k, m in itertools.product(range(n), repeat=4):
            v = h2[i, j, k, m]
            if v == 0:
                continue
            local = {}
            lorder = []
            for alpha, beta, gamma, delta in itertools.product(range(2), repeat=4):
                p1, ph1 = pauli_dot(a[i][alpha], a[k][beta])
                p2, ph2 = pauli_dot(p1, a[m][gamma])
                p3, ph3 = pauli_dot(p2, a[j][delta])
                coeff = v / 16 * (ph1 * ph2 * ph3) * (-1j) ** (alpha + beta) * (1j) ** (gamma + delta)
                if abs(coeff) > threshold:
                    if p3 not in local:
                        local[p3] = 0.0
                        lorder.append(p3)
                    local[p3] += coeff
            for p in lorder:
                if local[p] != 0:
                    add(local[p], p)
        chop_acc()
    return [(acc[p], masks_to_label(p[0], p[1], n)) for p in order]


def onee_to_spin(mo_onee: np.ndarray) -> np.ndarray:
    """qiskit/chemistry/qmolecule.py:431-453 (alpha == beta block-diagonal spin form)."""
    n = mo_onee.shape[0]
    blk = np.where(np.abs(mo_onee) > 1e-12, mo_onee, 0.0)
    out = np.zeros((2 * n, 2 * n))
    out[:n, :n] = blk
    out[n:, n:] = blk
    return out


def twoe_to_spin(mo_eri: np.ndarray) -> np.ndarray:
    """qiskit/chemistry/qmolecule.py:456-534: einsum 'ijkl->ljik' (:484) of the MO ERIs into the
    four same/opposite-spin blocks, then the -0.5 factor (:532)."""
    n = mo_eri.shape[0]
    ints = np.einsum('ijkl->ljik', mo_eri)
    out = np.zeros((2 * n,) * 4)
    for p, q, r, s in itertools.product(range(n), repeat=4):
        v = ints[p, q, r, s]
        if not abs(v) > 1e-12:
            continue
        out[p, q, r, s] = v                    # aa aa
        out[p + n, q + n, r + n, s + n] = v    # bb bb
        out[p, q + n, r + n, s] = v            # a b b a
        out[p + n, q, r, s + n] = v            # b a a b
    return -0.5 * out
