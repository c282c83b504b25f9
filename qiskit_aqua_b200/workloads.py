"""Seeded synthetic operators for the BASELINE.json configs, as packed
(x_mask, z_mask, ypow, coeff) arrays.  All RNG is np.random.default_rng(seed), matching
aqua_globals.random (reference qiskit/aqua/aqua_globals.py:89-96)."""
import itertools

import numpy as np


def _finish(n, terms):
    """terms: dict (x, z) -> real/complex coefficient (duplicates already merged)."""
    xs = np.array([k[0] for k in terms], dtype=np.uint64)
    zs = np.array([k[1] for k in terms], dtype=np.uint64)
    ys = np.array([bin(k[0] & k[1]).count('1') & 3 for k in terms], dtype=np.uint8)
    cs = np.array(list(terms.values()), dtype=np.complex128)
    return n, xs, zs, ys, cs


def _add(terms, x, z, c):
    terms[(x, z)] = terms.get((x, z), 0.0) + c


def two_local_random(n=24, n_terms=2000, seed=24):
    """C2: distinct 2-local Paulis, pair (i<j) uniform, each site in {X,Y,Z}, c ~ N(0,1) real."""
    rng = np.random.default_rng(seed)
    terms = {}
    while len(terms) < n_terms:
        i, j = sorted(rng.choice(n, size=2, replace=False).tolist())
        x = z = 0
        for q in (i, j):
            p = rng.integers(0, 3)
            if p in (0, 1):
                x |= 1 << q
            if p in (1, 2):
                z |= 1 << q
        c = rng.standard_normal()
        if (x, z) not in terms:
            terms[(x, z)] = c
    return _finish(n, terms)


def ising_heisenberg(n=30, n_triples=8665, seed=30, J=1.0, h=0.7):
    """C4: -J sum_{i<j nn} Z_i Z_{i+1} - h sum X_i + sum_{i<j} J_ij (XX + YY + ZZ) over all pairs
    + a seeded 3-local part sum K_ijk Z_i X_j Z_k over `n_triples` distinct random triples.
    n=30: 3*435 + 30 + 8665 = 10 000 terms, G_x = 1 + 30 + 435 = 466 distinct x-masks."""
    rng = np.random.default_rng(seed)
    terms = {}
    for i in range(n - 1):
        _add(terms, 0, (1 << i) | (1 << (i + 1)), -J)
    for i in range(n):
        _add(terms, 1 << i, 0, -h)
    for i, j in itertools.combinations(range(n), 2):
        jij = rng.standard_normal() / n
        m = (1 << i) | (1 << j)
        _add(terms, m, 0, jij)       # XX
        _add(terms, m, m, jij)       # YY
        _add(terms, 0, m, jij)       # ZZ
    seen = set()
    max_triples = n * (n - 1) * (n - 2) // 2
    n_triples = min(n_triples, max_triples)
    while len(seen) < n_triples:
        j = int(rng.integers(0, n))
        i, k = sorted(rng.choice([q for q in range(n) if q != j], size=2, replace=False).tolist())
        if (i, j, k) in seen:
            continue
        seen.add((i, j, k))
        _add(terms, 1 << j, (1 << i) | (1 << k), rng.standard_normal() / n)
    return _finish(n, terms)


def random_weighted_paulis(n=34, n_terms=10000, max_weight=4, seed=34):
    """C5: Paulis uniform over {I,X,Y,Z}^n restricted to weight 1..max_weight, c ~ N(0,1) real."""
    rng = np.random.default_rng(seed)
    terms = {}
    while len(terms) < n_terms:
        w = int(rng.integers(1, max_weight + 1))
        qs = rng.choice(n, size=w, replace=False)
        x = z = 0
        for q in qs.tolist():
            p = rng.integers(0, 3)
            if p in (0, 1):
                x |= 1 << q
            if p in (1, 2):
                z |= 1 << q
        if (x, z) not in terms:
            terms[(x, z)] = rng.standard_normal()
    return _finish(n, terms)


def random_dense_paulis(n, n_terms, seed, complex_coeffs=False):
    """Unstructured test operator: uniformly random labels (all weights), optional complex c."""
    rng = np.random.default_rng(seed)
    terms = {}
    while len(terms) < min(n_terms, 4 ** n):
        x = int(rng.integers(0, 1 << n, dtype=np.uint64)) if n < 63 else int(rng.integers(0, 1 << 62))
        z = int(rng.integers(0, 1 << n, dtype=np.uint64)) if n < 63 else int(rng.integers(0, 1 << 62))
        c = rng.standard_normal() + (1j * rng.standard_normal() if complex_coeffs else 0.0)
        if (x, z) not in terms:
            terms[(x, z)] = c
    return _finish(n, terms)


def molecular_style(n=14, n_quads=170, seed=14):
    """C3: synthetic molecular-style Hamiltonian: FermionicOperator with a seeded random symmetric
    h1 and a sparse random h2 (chemists' index convention of the class, symmetrised so that H is
    Hermitian), Jordan-Wigner mapped (reference qiskit/chemistry/fermionic_operator.py:344-474).
    Returns (n, x, z, ypow, coeff) and the realised number of Pauli terms is len(x)."""
    from .applications import FermionicOperator
    rng = np.random.default_rng(seed)
    h1 = rng.standard_normal((n, n))
    h1 = 0.5 * (h1 + h1.T)
    h2 = np.zeros((n, n, n, n))
    for _ in range(n_quads):
        i, j, k, m = (int(v) for v in rng.integers(0, n, size=4))
        v = 0.5 * rng.standard_normal()
        for a, b, c, d in ((i, j, k, m), (j, i, m, k), (k, m, i, j), (m, k, j, i)):
            h2[a, b, c, d] += v
    op = FermionicOperator(h1, h2).mapping('jordan_wigner')
    terms = {(p.x_mask, p.z_mask): complex(w) for w, p in op.paulis}
    return _finish(n, terms)
