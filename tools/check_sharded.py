#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
Shards a statevector by its top log2(N) qubits, runs the NCCL-exchanging H.psi / <H> / Lanczos
of libpsum.so and compares with the CPU oracle on the gathered result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co                                   # noqa: E402
from qiskit_aqua_b200 import PauliSumEngine, fill_state, nccl_unique_id, workloads as W  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
    ok = True
    for name, (n, x, z, y, c), opts in [
            ('two_local_23q', W.two_local_random(23, 600, seed=23), {}),   # n_loc >= 20: chunked host path
            ('two_local_20q', W.two_local_random(20, 600, seed=20), {}),
            ('weight4_18q', W.random_weighted_paulis(18, 800, 4, seed=18), {}),
            ('dense_complex_14q', W.random_dense_paulis(14, 200, seed=3, complex_coeffs=True), {}),
            ('stream_12q', W.random_dense_paulis(12, 100, seed=4, complex_coeffs=True), {'kernel': 1})]:
        p_bits = world.bit_length() - 1
        N_loc = 1 << (n - p_bits)
        eng = PauliSumEngine(n, x, z, y, c, opts)
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.shard(rank, world, uid[0])
        psi = fill_state(N_loc, seed=11, first_index=rank * N_loc)
        recv = torch.empty(max(2, world - 1) * N_loc, dtype=torch.complex128, device='cuda')
        for rcv in (recv, recv[:2 * N_loc], recv[:N_loc]):   # ring of world-1, 2 and 1 shard buffers
            yv, ex = eng.apply_sharded(psi, recv=rcv, want_expect=True)
            _, ex2 = eng.apply_sharded(psi, recv=rcv, want_expect=True, store=False)
            host = psi.cpu().pin_memory()
            ex3 = eng.expect_sharded_host(host, scratch=torch.empty_like(psi), recv=rcv)
            gathered = [torch.empty_like(yv) for _ in range(world)]
            dist.all_gather(gathered, yv)
            if rank == 0:
                full = co.counter_state(11, 0, 1 << n)
                ref = co.apply(n, x, z, y, c, full)
                got = torch.cat(gathered).cpu().numpy()
                err = np.abs(got - ref).max() / np.abs(ref).max()
                eref = np.vdot(full, ref)
                eerr = abs(ex - eref) / abs(eref)
                eerr2 = max(abs(ex2 - eref), abs(ex3 - eref)) / abs(eref)
                good = err < 1e-10 and eerr < 1e-10 and eerr2 < 1e-10
                ok &= good
                print('%-18s world=%d parts=%d passes=%d  y err %.2e  <H> err %.2e / %.2e  %s' %
                      (name, world, eng.info()['n_global_parts'], eng.info()['n_passes'], err, eerr, eerr2,
                       'OK' if good else 'FAIL'), flush=True)
        if eng.info()['is_hermitian']:
            r = eng.lanczos(k=2, tol=1e-12, want_residuals=True)
            single = None
            if rank == 0:
                e1 = PauliSumEngine(n, x, z, y, c, opts)
                single = e1.lanczos(k=2, tol=1e-12)['evals']
                good = np.abs(r['evals'] - single).max() < 1e-9 * np.abs(single).max() and r['resid'].max() < 1e-6
                ok &= bool(good)
                print('%-18s sharded lanczos %s vs single %s resid %.1e %s' %
                      (name, r['evals'], single, r['resid'].max(), 'OK' if good else 'FAIL'), flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == '__main__':
    main()
