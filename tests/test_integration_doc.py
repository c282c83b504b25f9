"""The reference-side ctypes stub printed in INTEGRATION.md is executable as written: it loads the
library, packs a list of [coeff, Pauli] terms and gets a handle whose info matches (no device)."""
import ctypes as C
import os
import re

import numpy as np

import qiskit_aqua_b200 as q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs():
    text = open(os.path.join(ROOT, 'INTEGRATION.md'), encoding='utf8').read()
    block = re.search(r'```python\n(# qiskit/aqua/operators/legacy/_psum\.py.*?)```', text, re.S).group(1)
    so = os.path.join(ROOT, 'qiskit_aqua_b200', 'libpsum.so')
    block = block.replace("C.CDLL('libpsum.so')", 'C.CDLL(%r)' % so)
    ns = {'AquaError': q.AquaError}
    exec(compile(block, 'INTEGRATION.md', 'exec'), ns)
    paulis = [[0.5, q.Pauli('XZ')], [-1.25 + 0.5j, q.Pauli('YY')], [2.0, q.Pauli('IZ')], [0.25, q.Pauli('XZ')]]
    h = ns['pack'](paulis)
    assert h.value
    lib = ns['_lib']
    lib.psum_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    info = (C.c_int64 * 15)()
    assert lib.psum_info(h, info, 15) == 0
    assert info[0] == 2 and info[1] == 3            # qubits, terms after merging the duplicate XZ
    lib.psum_destroy.argtypes = [C.c_void_p]
    assert lib.psum_destroy(h) == 0
    # the packing convention of the stub equals the package's own
    bits = 1 << np.arange(2, dtype=np.uint64)
    assert int((q.Pauli('XZ').x * bits).sum()) == q.Pauli('XZ').x_mask
    assert int((q.Pauli('XZ').z * bits).sum()) == q.Pauli('XZ').z_mask
