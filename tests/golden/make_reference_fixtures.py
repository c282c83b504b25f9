"""Runs the REFERENCE'S OWN code for the hot path in the build container and records its outputs
as golden fixtures (tests/golden/reference_run.json).

    python tests/golden/make_reference_fixtures.py        (needs /root/reference; CPU only)

qiskit-terra is absent, so tests/golden/terra_stub.py supplies a restated `Pauli` and dummies for
everything else; only the reference files on the path are executed for real:
  qiskit/aqua/operators/legacy/{weighted_pauli_operator,op_converter,matrix_operator,base_operator,common}.py
  qiskit/aqua/{aqua_error,aqua_globals,missing_optional_library_error}.py
  qiskit/optimization/applications/ising/{max_cut,common}.py
  qiskit/chemistry/fermionic_operator.py
  qiskit/aqua/operators/{operator_base,primitive_ops/{primitive_op,pauli_op},list_ops/{list_op,summed_op},
                         state_fns/*}.py  (opflow core, enough for to_opflow()/to_spmatrix())
  qiskit/aqua/algorithms/{eigen_solvers/numpy_eigen_solver,minimum_eigen_solvers/numpy_minimum_eigen_solver}.py
  qiskit/aqua/operators/{primitive_ops/matrix_op,list_ops/composed_op,expectations/matrix_expectation}.py
  (terra's Statevector / quantum_info.Operator are functional minimal stubs: data + dims)
The package __init__ files of the reference are NOT executed (they import the whole of terra);
synthetic parent packages give the relative imports something to hang on.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import terra_stub  # noqa: E402

REF = '/root/reference/qiskit'


def load_reference():
    terra_stub.install()
    import qiskit  # noqa: F401  (stub package)

    import importlib

    def synth(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
        parent, _, leaf = name.rpartition('.')
        setattr(sys.modules[parent], leaf, m)
        return m

    aqua = synth('qiskit.aqua', REF + '/aqua')
    ae = importlib.import_module('qiskit.aqua.aqua_error')
    ag = importlib.import_module('qiskit.aqua.aqua_globals')
    mo = importlib.import_module('qiskit.aqua.missing_optional_library_error')
    aqua.AquaError, aqua.aqua_globals, aqua.MissingOptionalLibraryError = ae.AquaError, ag.aqua_globals, mo.MissingOptionalLibraryError
    ops = synth('qiskit.aqua.operators', REF + '/aqua/operators')
    for sub in ('legacy', 'primitive_ops', 'list_ops', 'state_fns'):
        synth('qiskit.aqua.operators.' + sub, REF + '/aqua/operators/' + sub)
    # opflow core; operator_globals is synthetic (the real one builds gate circuits at import time)
    og = types.ModuleType('qiskit.aqua.operators.operator_globals')
    og.EVAL_SIG_DIGITS = 18
    sys.modules[og.__name__] = og
    ops.operator_globals = og
    ops.OperatorBase = importlib.import_module('qiskit.aqua.operators.operator_base').OperatorBase
    lb = importlib.import_module('qiskit.aqua.operators.legacy.base_operator')
    ops.LegacyBaseOperator = sys.modules['qiskit.aqua.operators.legacy'].LegacyBaseOperator = lb.LegacyBaseOperator
    ops.PrimitiveOp = importlib.import_module('qiskit.aqua.operators.primitive_ops.primitive_op').PrimitiveOp
    ops.PauliOp = importlib.import_module('qiskit.aqua.operators.primitive_ops.pauli_op').PauliOp
    ops.ListOp = importlib.import_module('qiskit.aqua.operators.list_ops.list_op').ListOp
    ops.SummedOp = importlib.import_module('qiskit.aqua.operators.list_ops.summed_op').SummedOp
    ops.StateFn = importlib.import_module('qiskit.aqua.operators.state_fns.state_fn').StateFn
    for nm in ('vector_state_fn', 'operator_state_fn', 'dict_state_fn', 'circuit_state_fn'):
        importlib.import_module('qiskit.aqua.operators.state_fns.' + nm)
    og.I = ops.I = ops.PauliOp(terra_stub.Pauli('I'))
    # the rest of what the MatrixExpectation chain touches
    for nm in ('composed_op', 'tensored_op'):
        importlib.import_module('qiskit.aqua.operators.list_ops.' + nm)
    ops.ComposedOp = sys.modules['qiskit.aqua.operators.list_ops.composed_op'].ComposedOp
    ops.TensoredOp = sys.modules['qiskit.aqua.operators.list_ops.tensored_op'].TensoredOp
    ops.MatrixOp = importlib.import_module('qiskit.aqua.operators.primitive_ops.matrix_op').MatrixOp
    ops.CircuitOp = importlib.import_module('qiskit.aqua.operators.primitive_ops.circuit_op').CircuitOp
    for cls, mod in (('VectorStateFn', 'vector_state_fn'), ('OperatorStateFn', 'operator_state_fn'),
                     ('DictStateFn', 'dict_state_fn'), ('CircuitStateFn', 'circuit_state_fn')):
        setattr(ops, cls, getattr(sys.modules['qiskit.aqua.operators.state_fns.' + mod], cls))
    lo = sys.modules['qiskit.aqua.operators.list_ops']
    lo.ListOp, lo.ComposedOp, lo.SummedOp, lo.TensoredOp = ops.ListOp, ops.ComposedOp, ops.SummedOp, ops.TensoredOp
    sfm = sys.modules['qiskit.aqua.operators.state_fns']
    for cls in ('StateFn', 'VectorStateFn', 'OperatorStateFn', 'DictStateFn', 'CircuitStateFn'):
        setattr(sfm, cls, getattr(ops, cls))
    pom = sys.modules['qiskit.aqua.operators.primitive_ops']
    pom.PrimitiveOp, pom.PauliOp, pom.MatrixOp = ops.PrimitiveOp, ops.PauliOp, ops.MatrixOp
    ops.Zero, ops.One = ops.StateFn('0'), ops.StateFn('1')
    cv = synth('qiskit.aqua.operators.converters', REF + '/aqua/operators/converters')
    cv.ConverterBase = importlib.import_module('qiskit.aqua.operators.converters.converter_base').ConverterBase
    synth('qiskit.aqua.operators.expectations', REF + '/aqua/operators/expectations')
    importlib.import_module('qiskit.aqua.operators.expectations.expectation_base')
    mexp = importlib.import_module('qiskit.aqua.operators.expectations.matrix_expectation')
    from qiskit.aqua.operators.legacy.weighted_pauli_operator import WeightedPauliOperator
    from qiskit.aqua.operators.legacy.matrix_operator import MatrixOperator
    from qiskit.aqua.operators.legacy import op_converter
    ops.WeightedPauliOperator = WeightedPauliOperator
    synth('qiskit.aqua.utils', REF + '/aqua/utils')
    algs = synth('qiskit.aqua.algorithms', REF + '/aqua/algorithms')
    algs.AlgorithmResult = importlib.import_module('qiskit.aqua.algorithms.algorithm_result').AlgorithmResult
    algs.ClassicalAlgorithm = importlib.import_module('qiskit.aqua.algorithms.classical_algorithm').ClassicalAlgorithm
    synth('qiskit.aqua.algorithms.eigen_solvers', REF + '/aqua/algorithms/eigen_solvers')
    synth('qiskit.aqua.algorithms.minimum_eigen_solvers', REF + '/aqua/algorithms/minimum_eigen_solvers')
    nes = importlib.import_module('qiskit.aqua.algorithms.eigen_solvers.numpy_eigen_solver')
    algs.NumPyEigensolver = sys.modules['qiskit.aqua.algorithms.eigen_solvers'].NumPyEigensolver = nes.NumPyEigensolver
    nmes = importlib.import_module('qiskit.aqua.algorithms.minimum_eigen_solvers.numpy_minimum_eigen_solver')
    synth('qiskit.optimization', REF + '/optimization')
    synth('qiskit.optimization.applications', REF + '/optimization/applications')
    synth('qiskit.optimization.applications.ising', REF + '/optimization/applications/ising')
    from qiskit.optimization.applications.ising import common, max_cut
    chem = synth('qiskit.chemistry', REF + '/chemistry')
    ce = importlib.import_module('qiskit.chemistry.qiskit_chemistry_error')
    chem.QiskitChemistryError = ce.QiskitChemistryError
    sys.modules.setdefault('retworkx', types.ModuleType('retworkx'))   # only bksf.py wants it; unused here
    from qiskit.chemistry.fermionic_operator import FermionicOperator
    return dict(WPO=WeightedPauliOperator, MatrixOperator=MatrixOperator, op_converter=op_converter,
                common=common, max_cut=max_cut, FermionicOperator=FermionicOperator, aqua_globals=ag.aqua_globals,
                NumPyEigensolver=nes.NumPyEigensolver, NumPyMinimumEigensolver=nmes.NumPyMinimumEigensolver,
                MatrixExpectation=mexp.MatrixExpectation, StateFn=ops.StateFn, ListOp=ops.ListOp, PauliOp=ops.PauliOp,
                AquaError=ae.AquaError)


def cplx(v):
    v = complex(v)
    return [v.real, v.imag]


def main():
    R = load_reference()
    WPO, Pauli = R['WPO'], terra_stub.Pauli
    out = {'note': 'outputs of the reference code itself (terra Pauli stubbed); see make_reference_fixtures.py',
           'operators': [], 'maxcut': [], 'fermionic': [], 'eigensolver': []}
    rng = np.random.default_rng(2024)
    for case in range(14):
        n = int(rng.integers(1, 7))
        T = int(rng.integers(1, 40))
        labels = [''.join(rng.choice(list('IXYZ')) for _ in range(n)) for _ in range(T)]
        if T > 3:                                    # force duplicates and an exact cancellation
            labels[1] = labels[0]
            labels[-1] = labels[2]
        weights = (rng.standard_normal(T) + (1j * rng.standard_normal(T) if case % 3 == 0 else 0)).astype(complex)
        if T > 3:
            weights[-1] = -weights[2]
        op = WPO(paulis=[[complex(w), Pauli(l)] for w, l in zip(weights, labels)])
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        rec = {'n': n, 'input_labels': labels, 'input_weights': [cplx(w) for w in weights],
               'simplified': [[cplx(w), p.to_label()] for w, p in op.paulis],
               'psi': [cplx(v) for v in psi]}
        if not op.is_empty():
            mean, std = op.evaluate_with_statevector(psi)
            rec['evaluate_with_statevector'] = cplx(mean)
            rec['std'] = std
            mat = R['op_converter'].to_matrix_operator(op)
            rec['matrix_operator_evaluate'] = cplx(mat.evaluate_with_statevector(psi)[0])
            hpsi = mat._matrix.dot(psi)
            rec['h_psi'] = [cplx(v) for v in hpsi]
            rec['to_dict'] = op.to_dict()
            # opflow chain: MatrixExpectation.convert(~StateFn(H) @ StateFn(psi)).eval()
            expr = ~R['StateFn'](op.to_opflow()) @ R['StateFn'](psi)
            rec['matrix_expectation'] = cplx(R['MatrixExpectation']().convert(expr).eval())
            rec['chop_0.5'] = [[cplx(w), p.to_label()] for w, p in op.chop(0.5, copy=True).paulis]
        else:
            try:
                op.evaluate_with_statevector(psi)
                rec['empty_raises'] = False
            except R['AquaError']:
                rec['empty_raises'] = True
        out['operators'].append(rec)
    # MaxCut generators (configs 1, G7)
    for n, seed, p in ((4, 8123179999, 0.5), (6, 99, 0.5), (16, 17, 0.5)):
        w = R['common'].random_graph(n, weight_range=10, edge_prob=p, negative_weight=True, seed=seed)
        op, shift = R['max_cut'].get_operator(w)
        out['maxcut'].append({'n': n, 'seed': seed, 'edge_prob': p, 'w': w.tolist(), 'shift': float(shift),
                              'paulis': [[cplx(wt), pl.to_label()] for wt, pl in op.paulis]})
    # Jordan-Wigner mapping (configs 3, G10): H2 integrals + a random sparse case
    chem = json.load(open(os.path.join(HERE, 'chemistry_integrals.json')))['G10_h2_fcidump']
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import pauli_oracle as po
    h1 = po.onee_to_spin(np.array(chem['mo_onee']))
    h2 = po.twoe_to_spin(np.array(chem['mo_eri']))
    cases = [('h2_fcidump', h1, h2)]
    r2 = np.random.default_rng(7)
    g1 = r2.standard_normal((5, 5))
    g1 = g1 + g1.T
    g2 = np.zeros((5,) * 4)
    for _ in range(10):
        i, j, k, m = r2.integers(0, 5, size=4)
        g2[i, j, k, m] = r2.standard_normal()
    cases.append(('random5', g1, g2))
    for name, a, b in cases:
        op = R['FermionicOperator'](a, b).mapping('jordan_wigner')
        out['fermionic'].append({'name': name, 'h1': a.tolist(), 'h2_nonzero': [[int(v) for v in idx] + [float(b[idx])]
                                                                               for idx in zip(*np.nonzero(b))],
                                 'modes': int(a.shape[0]),
                                 'paulis': [[cplx(w), p.to_label()] for w, p in op.paulis]})
    # NumPyEigensolver / NumPyMinimumEigensolver: all three branches (diagonal sort, dense eig,
    # ARPACK eigs), aux operators, filters -- through the reference's own _run()
    def aux_list(vals):
        return None if vals is None else [[None if a is None else [float(a[0]), float(a[1])] for a in row] for row in vals]

    H2 = [(-1.052373245772859, 'II'), (0.39793742484318045, 'ZI'), (-0.39793742484318045, 'IZ'),
          (-0.01128010425623538, 'ZZ'), (0.18093119978423156, 'XX')]
    AUX = [[(2.0, 'II')], [(0.5, 'II'), (0.5, 'ZZ'), (0.5, 'YY'), (-0.5, 'XX')]]
    eig_cases = [('h2_k1', H2, 1, AUX, None), ('h2_k4', H2, 4, AUX, None), ('h2_k3_noaux', H2, 3, None, None),
                 ('h2_filter', H2, 2, AUX, -1.0)]
    r3 = np.random.default_rng(11)
    for n, k in ((3, 2), (4, 3), (5, 1), (6, 4)):
        labs = sorted({''.join(r3.choice(list('IXYZ')) for _ in range(n)) for _ in range(25)})
        terms = [(float(r3.standard_normal()), l) for l in labs]
        aux = [[(float(r3.standard_normal()), l) for l in labs[:5]], [(1.0, 'Z' * n)]]
        eig_cases.append(('random_n%d_k%d' % (n, k), terms, k, aux, None))
    for n, seed in ((4, 8123179999), (6, 99)):
        w = R['common'].random_graph(n, weight_range=10, edge_prob=0.5, negative_weight=True, seed=seed)
        op, _ = R['max_cut'].get_operator(w)
        eig_cases.append(('maxcut_n%d' % n, [(complex(wt).real, pl.to_label()) for wt, pl in op.paulis], 2,
                          [[(1.0, 'Z' + 'I' * (n - 1))]], None))
    for name, terms, k, aux, thr in eig_cases:
        op = WPO(paulis=[[w, Pauli(l)] for w, l in terms])
        aux_ops = None if aux is None else [WPO(paulis=[[w, Pauli(l)] for w, l in a]) for a in aux]
        fc = None if thr is None else (lambda x, v, a, t=thr: v >= t)
        res = R['NumPyEigensolver'](op, k=k, aux_operators=aux_ops, filter_criterion=fc).run()
        mres = R['NumPyMinimumEigensolver'](op, aux_operators=aux_ops, filter_criterion=fc).run()
        vecs = [np.asarray(sf.primitive.data if hasattr(sf.primitive, 'data') and isinstance(sf.primitive.data, np.ndarray)
                           else 0) for sf in res.eigenstates]
        out['eigensolver'].append({
            'name': name, 'terms': [[w, l] for w, l in terms], 'k': k,
            'aux': None if aux is None else [[[w, l] for w, l in a] for a in aux], 'filter_ge': thr,
            'eigenvalues': [cplx(v) for v in res.eigenvalues],
            'aux_operator_eigenvalues': aux_list(res.aux_operator_eigenvalues),
            'min_eigenvalue': None if mres.eigenvalue is None else cplx(mres.eigenvalue),
            'min_aux': None if mres.aux_operator_eigenvalues is None else
            [None if a is None else [float(a[0]), float(a[1])] for a in mres.aux_operator_eigenvalues]})
    # G5 table through the reference chain: <psi|P|psi> for P in X,Y,Z,I on |+>,|->,|0>,|1>,
    # evaluated as one ListOp of states (matrix_op.py:199-201 broadcast)
    sq = 1 / np.sqrt(2)
    states = [[sq, sq], [sq, -sq], [1, 0], [0, 1]]
    table = []
    for lab in 'XYZI':
        meas = ~R['StateFn'](R['PauliOp'](Pauli(lab)))
        expr = meas @ R['ListOp']([R['StateFn'](np.array(v, dtype=complex)) for v in states])
        vals = R['MatrixExpectation']().convert(expr).eval()
        table.append([cplx(v) for v in vals])
    out['matrix_expectation_table'] = {'paulis': 'XYZI', 'states': states, 'values': table}
    # coefficient semantics of the measurement chain: a ListOp front with a non-unit (complex) list
    # coefficient (operator_state_fn.py:199-203 folds it into each element BEFORE the evaluation),
    # element coefficients, a measurement coefficient, and VectorStateFn measurements with complex
    # coefficients (vector_state_fn.py:76-79 adjoint stores the conjugated vector, :190-192 plain dot)
    rng = np.random.default_rng(99)
    coeff_cases = []
    for lab, lcoef, mcoef in (('ZX', 2.0, 1.0), ('XY', 1 + 1j, 0.5), ('YZ', -0.5j, 2 - 1j)):
        vs = [rng.standard_normal(4) + 1j * rng.standard_normal(4) for _ in range(3)]
        ecs = [1.0, 0.5 - 0.25j, 2.0]
        meas = (~R['StateFn'](R['PauliOp'](Pauli(lab)))) * mcoef
        front = R['ListOp']([R['StateFn'](v) * ec for v, ec in zip(vs, ecs)], coeff=lcoef)
        vals = R['MatrixExpectation']().convert(meas @ front).eval()
        coeff_cases.append({'pauli': lab, 'list_coeff': cplx(lcoef), 'meas_coeff': cplx(mcoef),
                            'elem_coeffs': [cplx(e) for e in ecs], 'states': [[cplx(a) for a in v] for v in vs],
                            'values': [cplx(v) for v in vals]})
    out['listop_coeff_cases'] = coeff_cases
    vec_cases = []
    for c_bra, c_ket in ((1.0, 1.0), (0.5 + 2j, 1.0), (1 - 1j, -0.25j)):
        a = rng.standard_normal(4) + 1j * rng.standard_normal(4)
        b = rng.standard_normal(4) + 1j * rng.standard_normal(4)
        bra_adj = (R['StateFn'](a) * c_bra).adjoint()                 # ~(c a): stores conj(a), conj(c)
        bra_direct = R['StateFn'](a, coeff=c_bra, is_measurement=True)   # stored vector used as is
        ket = R['StateFn'](b) * c_ket
        vec_cases.append({'a': [cplx(v) for v in a], 'b': [cplx(v) for v in b], 'c_bra': cplx(c_bra), 'c_ket': cplx(c_ket),
                          'adjoint_eval': cplx(bra_adj.eval(ket)), 'direct_eval': cplx(bra_direct.eval(ket)),
                          'adjoint_to_matrix': [cplx(v) for v in np.asarray(bra_adj.to_matrix()).reshape(-1)]})
    out['vector_measurement_cases'] = vec_cases
    path = os.path.join(HERE, 'reference_run.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('wrote', path, os.path.getsize(path), 'bytes;', len(out['operators']), 'operator cases')


if __name__ == '__main__':
    main()
