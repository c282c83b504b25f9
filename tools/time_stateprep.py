#!/usr/bin/env python
"""State preparation time of a 30-qubit EfficientSU2(reps=3, linear) ansatz: fused blocks vs gate by gate,
next to the <H> it feeds (C4 operator).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from qiskit_aqua_b200 import EfficientSU2, PauliSumEngine        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
circ = EfficientSU2(n, reps=3, entanglement='linear')
theta = np.random.default_rng(0).uniform(-np.pi, np.pi, circ.num_parameters)
buf = torch.empty(1 << n, dtype=torch.complex128, device='cuda')


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_fused = timed(lambda: circ.run(theta, out=buf))
blocks = circ.last_block_count
a = buf.clone()
ms_plain = timed(lambda: circ.run_unfused(theta, out=buf), reps=1)
diff = float(torch.linalg.vector_norm(a - buf).item())
_, x, z, y, c = bench.workload(n, 'c4')
eng = PauliSumEngine(n, x, z, y, c)
ms_expect = timed(lambda: eng.expect(buf))
print(json.dumps({'ansatz': 'EfficientSU2(%d, reps=3, linear)' % n, 'gates': len(circ.gates), 'blocks': blocks,
                  'fused_ms': ms_fused, 'gate_by_gate_ms': ms_plain, 'diff_norm': diff, 'expect_ms': ms_expect}))
