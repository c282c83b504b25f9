#!/usr/bin/env python
"""Batched <psi_b|H|psi_b> (psum_expect_batch) against a loop of psum_expect at 24 qubits (C2-style 2-local
2 000-term operator and the C4 family).  Prints one JSON line per operator."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_aqua_b200 import PauliSumEngine, fill_state, workloads as W     # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n, B = 24, 8
for name, (nn, x, z, y, c) in (('C2 two_local 24q 2000 terms', W.two_local_random(24, 2000, seed=24)),
                               ('C4-style 24q', W.ising_heisenberg(24, n_triples=10000 - 3 * 276 - 24, seed=30))):
    eng = PauliSumEngine(n, x, z, y, c)
    states = torch.stack([fill_state(1 << n, seed=b + 1) for b in range(B)])
    states /= torch.linalg.vector_norm(states, dim=1, keepdim=True)
    ms_batch = timed(lambda: eng.expect_batch(states))
    ms_loop = timed(lambda: [eng.expect(states[b]) for b in range(B)])
    a = eng.expect_batch(states)
    b_ = np.array([eng.expect(states[b]) for b in range(B)])
    print(json.dumps({'operator': name, 'batch': B, 'batched_ms_per_state': ms_batch / B, 'looped_ms_per_state': ms_loop / B,
                      'speedup': ms_loop / ms_batch, 'max_abs_diff': float(np.abs(a - b_).max()),
                      'passes': eng.info()['n_passes']}))
