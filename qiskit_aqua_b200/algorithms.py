"""NumPyEigensolver / NumPyMinimumEigensolver drop-ins on the device Lanczos + diagonal kernels.

Mirrors qiskit/aqua/algorithms/eigen_solvers/numpy_eigen_solver.py:47-308 and
qiskit/aqua/algorithms/minimum_eigen_solvers/numpy_minimum_eigen_solver.py:29-125:
  * diagonal operators take the sort branch (:163-169) -> psum_diag_kmin
  * otherwise the k algebraically smallest eigenpairs (scipy ``eigs(k, which='SR')`` :175-176,
    dense ``np.linalg.eig`` :171-173 when k >= 2^n - 1) -> psum_lanczos (Hermitian sums only;
    complex label coefficients raise AquaError -- there is no CPU fallback)
  * aux operators: <v|A|v>, real part, |.| <= 1e-12 -> 0, returned as (value, 0) (:208-226)
"""
import numpy as np
import torch

from ._lib import ERR_NOT_HERMITIAN, PsumError
from .operators import (AquaError, LegacyBaseOperator, ListOp, StateFn, WeightedPauliOperator,
                        _PauliSumMixin, _terms_of)


class AlgorithmResult(dict):
    @property
    def data(self):
        return self


class EigensolverResult(AlgorithmResult):
    """eigen_solver.py:102-134."""

    eigenvalues = property(lambda s: s.get('eigenvalues'), lambda s, v: s.__setitem__('eigenvalues', v))
    eigenstates = property(lambda s: s.get('eigenstates'), lambda s, v: s.__setitem__('eigenstates', v))
    aux_operator_eigenvalues = property(lambda s: s.get('aux_operator_eigenvalues'),
                                        lambda s, v: s.__setitem__('aux_operator_eigenvalues', v))


class MinimumEigensolverResult(AlgorithmResult):
    """minimum_eigen_solver.py:106-138."""

    eigenvalue = property(lambda s: s.get('eigenvalue'), lambda s, v: s.__setitem__('eigenvalue', v))
    eigenstate = property(lambda s: s.get('eigenstate'), lambda s, v: s.__setitem__('eigenstate', v))
    aux_operator_eigenvalues = property(lambda s: s.get('aux_operator_eigenvalues'),
                                        lambda s, v: s.__setitem__('aux_operator_eigenvalues', v))


def _to_opflow(op):
    if isinstance(op, LegacyBaseOperator):
        return op.to_opflow()
    return op


class NumPyEigensolver:
    """k lowest eigenpairs of a Pauli-sum operator, computed on the GPU."""

    def __init__(self, operator=None, k=1, aux_operators=None, filter_criterion=None,
                 tol=1e-12, max_iter=5000, seed=1):
        if k < 1:
            raise ValueError('k must have value >= 1, was %s' % k)      # validate_min (:70)
        self._operator = None
        self._aux_operators = []
        self._in_k = k
        self._k = k
        self.operator = operator
        self.aux_operators = aux_operators
        self._filter_criterion = filter_criterion
        self._tol, self._max_iter, self._seed = tol, max_iter, seed
        self._ret = {}

    # -- properties (same names as the reference) ------------------------------------------------
    @property
    def operator(self):
        return self._operator

    @operator.setter
    def operator(self, operator):
        self._operator = _to_opflow(operator) if operator is not None else None
        self._check_set_k()

    @property
    def aux_operators(self):
        return self._aux_operators

    @aux_operators.setter
    def aux_operators(self, aux_operators):
        if aux_operators is None:
            aux_operators = []
        elif not isinstance(aux_operators, list):
            aux_operators = [aux_operators]
        self._aux_operators = [None if op is None else _to_opflow(op) for op in aux_operators]

    @property
    def k(self):
        return self._in_k

    @k.setter
    def k(self, k):
        if k < 1:
            raise ValueError('k must have value >= 1, was %s' % k)
        self._in_k = k
        self._check_set_k()

    @property
    def filter_criterion(self):
        return self._filter_criterion

    @filter_criterion.setter
    def filter_criterion(self, fc):
        self._filter_criterion = fc

    @classmethod
    def supports_aux_operators(cls):
        return True

    def _check_set_k(self):
        if self._operator is not None:
            self._k = min(self._in_k, 2 ** self._operator.num_qubits)

    # -- solve --------------------------------------------------------------------------------------
    # dense branch (k >= 2^n - 1, or more pairs than the Krylov basis holds, or a tiny non-Hermitian sum):
    # the matrix is written on the device and diagonalised there by the dense library routine, as the
    # reference does with np.linalg.eig (:171-173).  Bounded by the size of the dense matrix.
    DENSE_MAX_QUBITS = 13
    LANCZOS_MAX_K = 120
    ARNOLDI_MIN_QUBITS = 7      # non-Hermitian sums above this size: device Arnoldi instead of the dense matrix

    def _solve_dense(self, eng, n, k_keep):
        if n > self.DENSE_MAX_QUBITS:
            raise AquaError('NumPyEigensolver: %d eigenpairs of a %d-qubit operator need the dense branch, which is '
                            'limited to %d qubits on the device (the reference would build a %d x %d matrix)'
                            % (k_keep, n, self.DENSE_MAX_QUBITS, 1 << n, 1 << n))
        mat = eng.to_dense()
        if eng.info()['is_hermitian']:
            w, v = torch.linalg.eigh(mat)
            w = w.to(torch.complex128)
        else:
            w, v = torch.linalg.eig(mat)
        wn = w.cpu().numpy()
        order = np.argsort(wn)[:k_keep] if k_keep > 1 else np.array([int(np.argmin(wn.real))])   # :177-180
        sel = torch.from_numpy(order.astype(np.int64)).cuda()
        self._ret['eigvals'] = wn[order]
        self._ret['eigvecs'] = v[:, sel].t().contiguous()

    def _solve(self):
        eng = self._operator._engine_for()
        info = eng.info()
        n = info['n_qubits']
        N = 1 << n
        if info['is_diagonal']:
            if self._k * N * 16 > 8 * 2 ** 30:
                raise AquaError('NumPyEigensolver: %d one-hot eigenvectors of a %d-qubit diagonal operator do not fit '
                                '(the reference allocates the same 2^n x k array)' % (self._k, n))
            if info['is_hermitian'] and self._k <= 32:
                vals, idx = eng.diag_kmin(self._k)                   # :163-169, k fused min-sweeps
                idx_t = torch.from_numpy(idx.astype(np.int64)).cuda()
                vals = vals.astype(np.float64)
            else:
                # many values (filter_criterion asks for all 2^n), or complex weights: one device sort.
                # np.sort / np.argsort order: ascending value (complex: real, then imaginary part),
                # ties by lower index -> stable sorts
                if info['is_hermitian']:
                    d = eng.diag()
                    _, order = torch.sort(d, stable=True)
                    vals = d[order[:self._k]].cpu().numpy()
                else:
                    ones = torch.ones(N, dtype=torch.complex128, device='cuda')
                    d = eng.apply(ones)                              # diagonal operator: (H 1)[i] = H[i, i]
                    _, o1 = torch.sort(d.imag, stable=True)
                    _, o2 = torch.sort(d.real[o1], stable=True)
                    order = o1[o2]
                    vals = d[order[:self._k]].cpu().numpy()
                idx_t = order[:self._k]
            vecs = torch.zeros((self._k, N), dtype=torch.complex128, device='cuda')
            vecs[torch.arange(self._k, device='cuda'), idx_t] = 1.0
            self._ret['eigvals'] = vals
            self._ret['eigvecs'] = vecs
            return
        # reference quirk (:171-173): for k >= 2^n - 1 the dense np.linalg.eig branch returns ALL 2^n
        # eigenpairs (not truncated to k); aux operators are still evaluated for the first k only
        if self._k >= N - 1:
            return self._solve_dense(eng, n, N)
        if not info['is_hermitian'] and n > self.ARNOLDI_MIN_QUBITS and self._k <= self.LANCZOS_MAX_K:
            # eigs(which='SR') accepts complex weights (ARPACK's Arnoldi): Krylov-Schur Arnoldi on the
            # device matvec; eigenvalues ascending by (real, imaginary) part like eigval.argsort() (:177-180)
            r = eng.arnoldi(k=self._k, tol=self._tol, max_iter=self._max_iter, seed=self._seed)
            self._ret['eigvals'] = r['evals']
            self._ret['eigvecs'] = r['evecs']
            return
        if not info['is_hermitian'] or self._k > self.LANCZOS_MAX_K:
            # tiny non-Hermitian sums, or more pairs than a Krylov basis holds: the dense routine
            if n > self.DENSE_MAX_QUBITS:
                raise AquaError('NumPyEigensolver on the GPU: k = %d > %d needs the dense branch, limited to %d '
                                'qubits; no CPU fallback is provided' % (self._k, self.LANCZOS_MAX_K,
                                                                         self.DENSE_MAX_QUBITS))
            return self._solve_dense(eng, n, self._k)
        r = eng.lanczos(k=self._k, tol=self._tol, max_iter=self._max_iter, seed=self._seed)
        self._ret['eigvals'] = r['evals'].astype(np.complex128)      # eigs returns complex values
        self._ret['eigvecs'] = r['evecs']

    def _eval_aux_operators(self, wavefn, threshold=1e-12):
        """:208-226 -- every aux operator against the eigenvector that is already on the device; all
        their passes are queued back to back and read with one synchronisation (psum_expect_multi)."""
        from .engine import expect_multi
        values = [None] * len(self._aux_operators)
        live = []
        for j, operator in enumerate(self._aux_operators):
            if operator is None:
                continue
            values[j] = (0.0, 0)
            terms = _terms_of(operator) if isinstance(operator, (_PauliSumMixin, WeightedPauliOperator)) else None
            if terms and any(w != 0 for w, _ in terms):
                live.append((j, operator._engine_for()))
        if live:
            res = expect_multi([e for _, e in live], wavefn)
            for (j, _), value in zip(live, res):
                values[j] = (value.real if abs(value.real) > threshold else 0.0, 0)
        return np.array(values, dtype=object)

    def _get_energies(self):
        self._ret['energies'] = np.array([np.real(v) for v in self._ret['eigvals']])
        if self._aux_operators:
            self._ret['aux_ops'] = [self._eval_aux_operators(self._ret['eigvecs'][i])
                                    for i in range(min(self._k, len(self._ret['eigvals'])))]

    def compute_eigenvalues(self, operator=None, aux_operators=None):
        if operator is not None:
            self.operator = operator
        if aux_operators is not None:
            self.aux_operators = aux_operators
        return self._run()

    def run(self):
        return self._run()

    def _run(self):
        if self._operator is None:
            raise AquaError("Operator was never provided")
        k_orig = self._k
        if self._filter_criterion:
            self._k = 2 ** self._operator.num_qubits                   # :249-251
        self._ret = {}
        self._solve()
        self._get_energies()
        if self._filter_criterion:
            keep = []
            for i in range(len(self._ret['eigvals'])):
                vec = self._ret['eigvecs'][i].cpu().numpy()
                aux = self._ret['aux_ops'][i] if 'aux_ops' in self._ret else None
                if self._filter_criterion(vec, self._ret['eigvals'][i], aux):
                    keep.append(i)
                if len(keep) == k_orig:
                    break
            sel = torch.tensor(keep, dtype=torch.long, device='cuda')
            self._ret['eigvecs'] = self._ret['eigvecs'][sel] if keep else self._ret['eigvecs'][:0]
            self._ret['eigvals'] = np.array([self._ret['eigvals'][i] for i in keep])
            self._ret['energies'] = np.array([self._ret['energies'][i] for i in keep])
            if 'aux_ops' in self._ret:
                self._ret['aux_ops'] = [self._ret['aux_ops'][i] for i in keep]
            self._k = k_orig
        if len(self._ret['eigvals']) > 0:
            self._ret['energy'] = float(np.real(self._ret['eigvals'][0]))
            self._ret['wavefunction'] = self._ret['eigvecs']
        else:
            self._ret['energy'] = None
            self._ret['wavefunction'] = None
        result = EigensolverResult()
        result.eigenvalues = self._ret['eigvals']
        result.eigenstates = ListOp([StateFn(vec) for vec in self._ret['eigvecs']])
        if 'aux_ops' in self._ret:
            result.aux_operator_eigenvalues = self._ret['aux_ops']
        return result


class NumPyMinimumEigensolver:
    """numpy_minimum_eigen_solver.py:29-125 (wraps NumPyEigensolver with k = 1)."""

    def __init__(self, operator=None, aux_operators=None, filter_criterion=None, **kwargs):
        self._ces = NumPyEigensolver(operator=operator, k=1, aux_operators=aux_operators,
                                     filter_criterion=filter_criterion, **kwargs)
        self._ret = {}

    operator = property(lambda s: s._ces.operator, lambda s, v: setattr(s._ces, 'operator', v))
    aux_operators = property(lambda s: s._ces.aux_operators, lambda s, v: setattr(s._ces, 'aux_operators', v))
    filter_criterion = property(lambda s: s._ces.filter_criterion,
                                lambda s, v: setattr(s._ces, 'filter_criterion', v))

    @classmethod
    def supports_aux_operators(cls):
        return True

    def compute_minimum_eigenvalue(self, operator=None, aux_operators=None):
        if operator is not None:
            self.operator = operator
        if aux_operators is not None:
            self.aux_operators = aux_operators
        return self._run()

    def run(self):
        return self._run()

    def _run(self):
        result_ces = self._ces.run()
        self._ret = self._ces._ret
        result = MinimumEigensolverResult()
        if len(result_ces.eigenvalues) > 0:
            result.eigenvalue = result_ces.eigenvalues[0]
            result.eigenstate = result_ces.eigenstates[0]
            if result_ces.aux_operator_eigenvalues is not None and len(result_ces.aux_operator_eigenvalues) > 0:
                result.aux_operator_eigenvalues = result_ces.aux_operator_eigenvalues[0]
        else:
            result.eigenvalue = None
            result.eigenstate = None
            result.aux_operator_eigenvalues = None
        return result
