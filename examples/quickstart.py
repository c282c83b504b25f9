#!/usr/bin/env python
"""Drop-in usage on a B200 (needs the built libpsum.so): the reference's entry points, unchanged.

    python examples/quickstart.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_aqua_b200 import (EfficientSU2, MatrixExpectation, NumPyEigensolver, NumPyMinimumEigensolver,  # noqa: E402
                              StateFn, VQE, WeightedPauliOperator, X, Z, max_cut, random_graph)

# 1. legacy operator, exactly as in test/aqua/test_numpy_eigen_solver.py
h2 = WeightedPauliOperator.from_dict({'paulis': [
    {"coeff": {"imag": 0.0, "real": -1.052373245772859}, "label": "II"},
    {"coeff": {"imag": 0.0, "real": 0.39793742484318045}, "label": "ZI"},
    {"coeff": {"imag": 0.0, "real": -0.39793742484318045}, "label": "IZ"},
    {"coeff": {"imag": 0.0, "real": -0.01128010425623538}, "label": "ZZ"},
    {"coeff": {"imag": 0.0, "real": 0.18093119978423156}, "label": "XX"}]})
print('spectrum      ', np.real(NumPyEigensolver(h2, k=4).run().eigenvalues))

# 2. <psi|H|psi> for a host state vector (copied to the device, no matrix is built)
psi = np.random.default_rng(0).standard_normal(4) + 0j
psi /= np.linalg.norm(psi)
print('<psi|H|psi>   ', h2.evaluate_with_statevector(psi))

# 3. opflow measurement
print('opflow        ', MatrixExpectation().convert(~StateFn(0.5 * (X ^ X) + (Z ^ Z)) @ StateFn(psi)).eval())

# 4. MaxCut through the diagonal branch
w = random_graph(16, weight_range=10, edge_prob=0.5, seed=17)
op, shift = max_cut.get_operator(w)
print('max-cut energy', NumPyMinimumEigensolver(op).run().eigenvalue.real + shift)

# 5. a VQE run whose ansatz state never leaves the device
print('VQE           ', VQE(h2, EfficientSU2(2, reps=2, entanglement='linear'), seed=1).run()['optimal_value'])
