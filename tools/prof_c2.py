import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qiskit_aqua_b200 import PauliSumEngine, fill_state, workloads as W
n, x, z, y, c = W.two_local_random(24, 2000, seed=24)
eng = PauliSumEngine(n, x, z, y, c)
psi = fill_state(1 << n, seed=3)
psi /= torch.linalg.vector_norm(psi)
print(eng.expect(psi))
yv = torch.empty_like(psi)
print(eng.apply(psi, out=yv, want_expect=True)[1])
torch.cuda.synchronize()
