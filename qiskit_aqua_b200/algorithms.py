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
    def _solve(self):
        eng = self._operator._engine_for()
        info = eng.info()
        n = info['n_qubits']
        N = 1 << n
        if info['is_diagonal']:
            vals, idx = eng.diag_kmin(self._k)                       # :163-169
            vecs = torch.zeros((self._k, N), dtype=torch.complex128, device='cuda')
            vecs[torch.arange(self._k, device='cuda'), torch.from_numpy(idx.astype(np.int64)).cuda()] = 1.0
            self._ret['eigvals'] = vals.astype(np.float64)
            self._ret['eigvecs'] = vecs
            return
        # reference quirk (:171-173): for k >= 2^n - 1 the dense np.linalg.eig branch returns ALL 2^n
        # eigenpairs (not truncated to k); aux operators are still evaluated for the first k only
        k_eff = N if self._k >= N - 1 else self._k
        try:
            r = eng.lanczos(k=k_eff, tol=self._tol, max_iter=self._max_iter, seed=self._seed)
        except PsumError as ex:
            if ex.code == ERR_NOT_HERMITIAN:
                raise AquaError('NumPyEigensolver on the GPU needs a Hermitian Pauli sum '
                                '(real coefficients); no CPU fallback is provided') from ex
            raise
        self._ret['eigvals'] = r['evals'].astype(np.complex128)      # eigs returns complex values
        self._ret['eigvecs'] = r['evecs']

    def _eval_aux_operators(self, wavefn, threshold=1e-12):
        values = []
        for operator in self._aux_operators:
            if operator is None:
                values.append(None)
                continue
            value = 0.0
            terms = _terms_of(operator) if isinstance(operator, (_PauliSumMixin, WeightedPauliOperator)) else None
            if terms and any(w != 0 for w, _ in terms):
                value = operator._engine_for().expect(wavefn)
                value = value.real if abs(value.real) > threshold else 0.0
            values.append((value, 0))
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
