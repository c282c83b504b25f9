"""qiskit_aqua_b200 -- B200-native drop-in for Aqua's Pauli-sum H.psi / <psi|H|psi> hot path.

Only the one hot path of qiskit-aqua 0.10.0 is provided; the compute runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/psum.h (libpsum.so).
"""
from ._lib import PsumError, lib  # noqa: F401
from .engine import PauliSumEngine, fill_state, nccl_unique_id, pack_masks  # noqa: F401

__version__ = '0.1.0'
from .operators import (AquaError, ComposedOp, I, ListOp, MatrixExpectation, OperatorStateFn,  # noqa: F401,E402
                        Pauli, PauliOp, StateFn, SummedOp, VectorStateFn, WeightedPauliOperator,
                        X, Y, Z)
from .algorithms import (EigensolverResult, MinimumEigensolverResult, NumPyEigensolver,  # noqa: F401,E402
                         NumPyMinimumEigensolver)
from .aqua_globals import aqua_globals  # noqa: F401,E402
from .applications import FermionicOperator  # noqa: F401,E402
from .ising import (clique, exact_cover, graph_partition, knapsack, max_cut, partition,  # noqa: F401,E402
                    random_graph, sample_most_likely, set_packing, stable_set, tsp, vehicle_routing,
                    vertex_cover)
from . import fcidump, ising, tapering  # noqa: F401,E402
from .tapering import Z2Symmetries  # noqa: F401,E402
from .legacy import MatrixOperator, op_converter  # noqa: F401,E402
from .fcidump import FCIDumpDriver, QiskitChemistryError, QMolecule  # noqa: F401,E402
from .circuits import (VQE, CircuitStateFn, EfficientSU2, RealAmplitudes, StatevectorCircuit,  # noqa: F401,E402
                       energy_evaluation)
