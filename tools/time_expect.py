import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from qiskit_aqua_b200 import PauliSumEngine, fill_state, workloads as W
from qiskit_aqua_b200.engine import vec_dot, vec_scale
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
pairs = n * (n - 1) // 2
_, x, z, y, c = W.ising_heisenberg(n, n_triples=max(0, 10000 - 3 * pairs - n), seed=30)
psi = fill_state(1 << n, seed=30)
vec_scale(psi, 1.0 / np.sqrt(vec_dot(psi, psi).real))
for hp in (0, 1):
    eng = PauliSumEngine(n, x, z, y, c)
    eng.set_option('herm_pairing', hp)
    e = eng.expect(psi)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): eng.expect(psi)
    ev0.record()
    for _ in range(3): e = eng.expect(psi)
    ev1.record(); torch.cuda.synchronize()
    print('herm_pairing=%d  <H>=%.15e %+.2ej   %.1f ms per expectation' % (hp, e.real, e.imag, ev0.elapsed_time(ev1) / 3))
