"""Meta-path stub for the absent qiskit-terra, used ONLY by make_reference_fixtures.py in the build
container (where /root/reference exists) to execute the reference's OWN Aqua code for the hot
path: WeightedPauliOperator (simplify / chop / from_dict / evaluate_with_statevector),
op_converter.to_matrix_operator, MatrixOperator, max_cut / common.random_graph and
FermionicOperator.mapping('jordan_wigner').

Only `Pauli` is a real implementation here -- terra's published convention restated (label
big-endian, (z, x) arrays little-endian, to_matrix rows i -> column i^x with
(-i)^{nY} (-1)^{popc(i&z)}, dot() returning the product with its phase as a power of -i);
every other terra name resolves to a permissive dummy so that module-level imports succeed.
Test infrastructure; never imported by the product."""
import importlib.abc, importlib.machinery, sys, types
import numpy as np
import scipy.sparse as sp

REF = '/root/reference'


class _Dummy:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Dummy()
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'): raise AttributeError(name)
        return _Dummy()
    def __iter__(self): return iter(())
    def __len__(self): return 0
    def __getitem__(self, k): return _Dummy()
    def __bool__(self): return False
    def __index__(self): return 1
    def __int__(self): return 1
    def __float__(self): return 0.0
    def __add__(self, o): return _Dummy()
    __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __matmul__ = __xor__ = __add__
    def __eq__(self, o): return False
    def __hash__(self): return 0
    def __enter__(self): return self
    def __exit__(self, *a): return False


class Pauli:
    """terra 0.18 Pauli restated: label big-endian, (z, x) arrays little-endian,
    to_matrix(sparse) rows i -> col i^x with (-i)^{nY} (-1)^{popc(i&z)}."""
    def __init__(self, data=None, x=None, z=None, label=None):
        if label is not None: data = label
        if isinstance(data, Pauli):
            self.z, self.x = data.z.copy(), data.x.copy(); return
        if isinstance(data, str):
            n = len(data)
            self.z = np.zeros(n, dtype=bool); self.x = np.zeros(n, dtype=bool)
            for pos, ch in enumerate(data):
                q = n - 1 - pos
                if ch in 'XY': self.x[q] = True
                if ch in 'ZY': self.z[q] = True
                if ch not in 'IXYZ': raise ValueError(ch)
            return
        if data is not None: z, x = data
        self.z = np.asarray(z).astype(bool); self.x = np.asarray(x).astype(bool)
    @classmethod
    def from_label(cls, label): return cls(label)
    @property
    def num_qubits(self): return len(self.z)
    def __len__(self): return len(self.z)
    def to_label(self):
        return ''.join('IXZY'[int(self.x[q]) + 2 * int(self.z[q])] for q in range(len(self.z) - 1, -1, -1))
    def __str__(self): return self.to_label()
    def __repr__(self): return "Pauli('%s')" % self.to_label()
    def __eq__(self, o): return isinstance(o, Pauli) and np.array_equal(self.z, o.z) and np.array_equal(self.x, o.x)
    def __hash__(self): return hash(self.to_label())
    def to_matrix(self, sparse=False):
        n = len(self.z); dim = 1 << n
        xm = sum(1 << q for q in range(n) if self.x[q]); zm = sum(1 << q for q in range(n) if self.z[q])
        rows = np.arange(dim, dtype=np.int64); cols = rows ^ xm
        par = np.zeros(dim, dtype=np.int64); v = rows & zm
        while v.any(): par ^= v & 1; v >>= 1
        data = ((-1j) ** bin(xm & zm).count('1')) * (1.0 - 2.0 * par).astype(complex)
        m = sp.csr_matrix((data, cols, np.arange(dim + 1)), shape=(dim, dim))
        return m if sparse else m.toarray()
    def to_spmatrix(self): return self.to_matrix(sparse=True)
    # ---- product with phase, as used by qiskit/chemistry/fermionic_operator.py:418-474 ----------
    phase = 0                      # label Pauli times (-i)^phase

    def dot(self, other):
        """self . other (matrix product) = (-i)^phase * label Pauli."""
        ax = sum(1 << q for q in range(len(self.x)) if self.x[q]); az = sum(1 << q for q in range(len(self.z)) if self.z[q])
        bx = sum(1 << q for q in range(len(other.x)) if other.x[q]); bz = sum(1 << q for q in range(len(other.z)) if other.z[q])
        cx, cz = ax ^ bx, az ^ bz
        e = (bin(ax & az).count('1') + bin(bx & bz).count('1') - bin(cx & cz).count('1')
             + 2 * bin(az & bx).count('1')) % 4            # product = i^e * label
        out = Pauli((self.z ^ other.z, self.x ^ other.x))
        out.phase = (-e + getattr(self, 'phase', 0) + getattr(other, 'phase', 0)) % 4
        return out

    def expand(self, other):       # other (x) self: `other` on the higher qubits
        return Pauli((np.concatenate([self.z, other.z]), np.concatenate([self.x, other.x])))

    def tensor(self, other):       # self (x) other: `self` on the higher qubits
        return Pauli((np.concatenate([other.z, self.z]), np.concatenate([other.x, self.x])))

    def __getitem__(self, item):   # p[:] drops the phase (terra: a fresh label Pauli)
        return Pauli((self.z[item], self.x[item]))


class Statevector:
    """Functional minimum of terra's Statevector (what VectorStateFn touches)."""
    def __init__(self, data, dims=None):
        self.data = np.asarray(data.data if isinstance(data, Statevector) else data, dtype=complex).reshape(-1)
    @property
    def num_qubits(self): return int(round(np.log2(self.data.shape[0])))
    @property
    def dim(self): return self.data.shape[0]
    def dims(self): return (2,) * self.num_qubits
    def conjugate(self): return Statevector(np.conj(self.data))
    def tensor(self, other): return Statevector(np.kron(self.data, other.data))
    def __add__(self, other): return Statevector(self.data + other.data)
    def __mul__(self, c): return Statevector(self.data * c)
    __rmul__ = __mul__
    def __eq__(self, other): return isinstance(other, Statevector) and np.allclose(self.data, other.data)
    def __hash__(self): return 0
    def to_dict(self): return {format(i, '0%db' % self.num_qubits): v for i, v in enumerate(self.data) if v != 0}


class Operator:
    """Functional minimum of terra's quantum_info.Operator (what MatrixOp touches)."""
    def __init__(self, data, input_dims=None, output_dims=None):
        self.data = np.asarray(data.data if isinstance(data, Operator) else data, dtype=complex)
    @property
    def num_qubits(self): return int(round(np.log2(self.data.shape[0])))
    @property
    def dim(self): return self.data.shape
    def input_dims(self): return (2,) * int(round(np.log2(self.data.shape[1])))
    def output_dims(self): return (2,) * int(round(np.log2(self.data.shape[0])))
    def adjoint(self): return Operator(self.data.conj().T)
    def conjugate(self): return Operator(self.data.conj())
    def transpose(self): return Operator(self.data.T)
    def tensor(self, other): return Operator(np.kron(self.data, other.data))
    def compose(self, other, front=False): return Operator(self.data @ other.data if front else other.data @ self.data)
    def dot(self, other): return Operator(self.data @ other.data)
    def __add__(self, other): return Operator(self.data + other.data)
    def __mul__(self, c): return Operator(self.data * c)
    __rmul__ = __mul__
    def __eq__(self, other): return isinstance(other, Operator) and np.allclose(self.data, other.data)
    def __hash__(self): return 0


SPECIAL = {
    ('qiskit.util', 'local_hardware_info'): lambda: {'cpus': 8, 'memory': 64, 'os': 'Linux'},
    ('qiskit.quantum_info', 'Pauli'): Pauli,
    ('qiskit.quantum_info.operators', 'Pauli'): Pauli,
}


def _parallel_map(task, values, task_args=(), task_kwargs=None, num_processes=1):
    return [task(v, *task_args, **(task_kwargs or {})) for v in values]


SPECIAL[('qiskit.tools', 'parallel_map')] = _parallel_map
SPECIAL[('qiskit.tools.parallel', 'parallel_map')] = _parallel_map


_CLASSES = {}


def get_class(name):
    """One stub class per terra class NAME (shared by all stub modules); *Gate classes derive from
    Instruction so that isinstance checks in the reference keep working."""
    if name in _CLASSES:
        return _CLASSES[name]
    bases = (object,)
    if name.endswith('Gate') and name != 'Gate':
        bases = (get_class('Gate'),)
    elif name in ('Gate', 'Measure', 'Barrier', 'Reset'):
        bases = (get_class('Instruction'),)
    body = {'__init__': lambda self, *a, **k: None,
            '__getattr__': lambda self, n: (_ for _ in ()).throw(AttributeError(n)) if n.startswith('__') else _Dummy(),
            '__call__': lambda self, *a, **k: _Dummy(),
            '__len__': lambda self: 0, '__iter__': lambda self: iter(())}
    cls = type(name, bases, body)
    _CLASSES[name] = cls
    return cls


class StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'): raise AttributeError(name)
        key = (self.__name__, name)
        if key in SPECIAL:
            return SPECIAL[key]
        if name == 'Pauli':
            return Pauli
        if name == 'Statevector':
            return Statevector
        if name == 'Operator':
            return Operator
        if name == 'parallel_map':
            return _parallel_map
        if name[0].isupper() and not name.isupper():
            obj = get_class(name)
        else:
            obj = _Dummy()
        setattr(self, name, obj)
        return obj


class TerraStubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    REAL = ('qiskit.aqua', 'qiskit.chemistry', 'qiskit.optimization', 'qiskit.finance', 'qiskit.ml')
    def find_spec(self, fullname, path, target=None):
        if fullname == 'qiskit' or (fullname.startswith('qiskit.') and not fullname.startswith(self.REAL)):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None
    def create_module(self, spec):
        m = StubModule(spec.name)
        m.__path__ = [REF + '/qiskit'] if spec.name == 'qiskit' else []
        return m
    def exec_module(self, module):
        if module.__name__ == 'qiskit':
            module.__version__ = '0.18.0'


def install():
    sys.meta_path.insert(0, TerraStubFinder())
