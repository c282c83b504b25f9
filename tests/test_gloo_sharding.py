"""N > 1 host logic on CPU: two gloo ranks shard the statevector by its top qubit, exchange
shards with send/recv exactly as the NCCL path does (rank r <-> r ^ g per global x-part g), and
apply the per-part terms produced by the library's sharding plan.  The arithmetic of each part is
done by the oracle here (the CUDA kernels need a GPU); what is under test is the plan algebra
(global/local mask split, rank sign folding) and the exchange schedule."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import pauli_oracle as po
    from qiskit_aqua_b200 import PauliSumEngine, workloads as W
    _, x, z, y, c = W.random_weighted_paulis(n, 150, max_weight=3, seed=21)
    rng = np.random.default_rng(7)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    p_bits = world.bit_length() - 1
    n_loc = n - p_bits
    N_loc = 1 << n_loc
    mine = psi[rank * N_loc:(rank + 1) * N_loc].copy()
    eng = PauliSumEngine(n, x, z, y, c)
    eng.shard(rank, world)
    yloc = np.zeros(N_loc, dtype=np.complex128)
    for g in range(world):
        xl, zl, cl = eng.shard_terms(g)
        if g == 0:
            src = mine
        else:
            peer = rank ^ g
            send = torch.from_numpy(mine.copy())
            recv = torch.empty_like(send)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, send, peer),
                                           dist.P2POp(dist.irecv, recv, peer)])
            for r in reqs:
                r.wait()
            src = recv.numpy()
        if len(xl):
            yloc += po.mf_apply(n_loc, xl, zl, np.zeros(len(xl), np.uint8), cl, src)
    e = torch.tensor([np.vdot(mine, yloc).real, np.vdot(mine, yloc).imag], dtype=torch.float64)
    dist.all_reduce(e)
    ref = po.mf_apply(n, x, z, y, c, psi)
    err = np.abs(yloc - ref[rank * N_loc:(rank + 1) * N_loc]).max() / np.abs(ref).max()
    eref = np.vdot(psi, ref)
    ret[rank] = (float(err), abs(complex(e[0].item(), e[1].item()) - eref) / abs(eref))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_sharded_matvec_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + world
    mp.spawn(_worker, args=(world, port, 8, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, eerr = ret[rank]
        assert err < 1e-12 and eerr < 1e-10
