#!/usr/bin/env python
"""bench.py -- headline benchmark of the Pauli-sum H.psi / <psi|H|psi> hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): Pauli-sum H.psi throughput in G term-amplitudes/s on the 30-qubit,
10 000-term transverse-field Ising + Heisenberg (+3-local) Hamiltonian (config C4).  One "step" =
one y = H.psi with the fused <psi|H|psi> reduction over a statevector resident in HBM.  With N > 1
ranks (torchrun) the statevector has 30 + log2(N) qubits, sharded by its top qubits (weak scaling:
2^30 amplitudes per GPU), with the NCCL amplitude exchange inside the step.

The JSON line also carries
  e2e          the same metric through the public API (WeightedPauliOperator.
               evaluate_with_statevector) with the state in pinned HOST memory: H2D copy of psi and
               D2H of the scalar inside the timed region
  roofline     algorithmic bytes 16 N (G_x + 1) of one H.psi / measured kernel time vs the measured
               HBM copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes from the ncu capture
               under profiles/ (null when not captured for this build)
  cpu_baseline the oracle C restatement (oracle/psum_oracle.c, OpenMP) timed on the host cores on a
               bounded row-sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'pauli_sum_apply_throughput'
UNIT = 'G term-amps/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--qubits', type=int, default=0, help='override qubits per GPU (default 30)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--workload', default='c4', choices=['c4', 'c5'],
                    help='c4: Ising+Heisenberg+3-local (headline); c5: random weight<=4 Pauli sum')
    ap.add_argument('--tile-bits', type=int, default=0)
    ap.add_argument('--low-bits', type=int, default=0)
    ap.add_argument('--reg-bits', type=int, default=0)
    return ap.parse_args()


def workload(n_total, kind='c4'):
    from qiskit_aqua_b200 import workloads as W
    if kind == 'c5':
        return W.random_weighted_paulis(n_total, 10000, 4, seed=34)
    pairs = n_total * (n_total - 1) // 2
    triples = max(0, 10000 - 3 * pairs - n_total - 0)
    # the chain ZZ terms coincide with pair ZZ terms (merged), so T = 3*pairs + n + triples
    return W.ising_heisenberg(n_total, n_triples=triples, seed=30)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([p.strip() for p in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for j, nm in enumerate(names) if any(s[2 + j].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': reasons,
                'samples': len(self.samples)}


def cpu_sample(n_total, x, z, y, c, psi_host, budget_s=12.0, rows_per_call=1 << 18):
    """Times the oracle C matvec on successive row blocks of the same operator until budget_s."""
    from oracle import c_oracle as co
    co.use_all_cores()
    T = len(x)
    done_rows, t_acc = 0, 0.0
    co.apply(n_total, x, z, y, c, psi_host, 0, 1 << 12)          # warm the threads
    N = 1 << n_total
    while t_acc < budget_s and done_rows < N:
        r1 = min(N, done_rows + rows_per_call)
        t0 = time.perf_counter()
        co.apply(n_total, x, z, y, c, psi_host, done_rows, r1)
        t_acc += time.perf_counter() - t0
        done_rows = r1
    return {'value': T * done_rows / t_acc / 1e9, 'unit': UNIT, 'cores': co.threads(), 'kind': 'port',
            'sample': 'rows [0, %d) of the %d-qubit %d-term H.psi (all terms), oracle/psum_oracle.c '
                      'OpenMP, %.1f s' % (done_rows, n_total, T, t_acc)}


def run_reference(args):
    """Reference arm: the reference's CPU path (matrix-free C port of its arithmetic -- the literal
    CSR build cannot be assembled at 30 qubits) on all host cores; each step = one row sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import c_oracle as co
    co.use_all_cores()                  # torchrun sets OMP_NUM_THREADS=1 for its workers
    n_loc = args.qubits or 30
    n_total = n_loc + (args.gpus.bit_length() - 1)
    n_cpu = min(n_total, 30)           # host RAM bound for the state the CPU streams over
    _, x, z, y, c = workload(n_total, args.workload)
    T = len(x)
    psi = co.counter_state(30, 0, 1 << n_cpu)
    rows = min(1 << 20, 1 << n_cpu)
    xs, zs, ys, cs = x & np.uint64((1 << n_cpu) - 1), z & np.uint64((1 << n_cpu) - 1), y, c
    times = []
    for s in range(args.warmup + args.steps):
        r0 = (s * rows) % (1 << n_cpu)
        t0 = time.perf_counter()
        co.apply(n_cpu, xs, zs, ys, cs, psi, r0, r0 + rows)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = T * rows / (ms * 1e-3) / 1e9
    sample = ('%d rows x %d terms of the %d-qubit workload per step (state truncated to %d qubits on '
              'the host), oracle/psum_oracle.c OpenMP' % (rows, T, n_total, n_cpu))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C4 ising+heisenberg+3local %dq %d terms (row sample)' % (n_total, T),
                       'qubits': n_total, 'terms': T},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': co.threads(), 'kind': 'port',
                             'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from qiskit_aqua_b200 import PauliSumEngine, WeightedPauliOperator, fill_state, nccl_unique_id
    from qiskit_aqua_b200.engine import vec_dot, vec_scale

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    n_loc = args.qubits or 30
    p_bits = world.bit_length() - 1
    n_total = n_loc + p_bits
    N_loc = 1 << n_loc
    _, x, z, y, c = workload(n_total, args.workload)
    T = len(x)
    opts = {}
    if args.tile_bits:
        opts['tile_bits'] = args.tile_bits
    if args.low_bits:
        opts['low_bits'] = args.low_bits
    if args.reg_bits:
        opts['reg_bits'] = args.reg_bits
    eng = PauliSumEngine(n_total, x, z, y, c, opts)
    recv = None
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.shard(rank, world, uid[0])
    info = eng.info()
    gx_total = info['n_xgroups']

    psi = fill_state(N_loc, seed=30, first_index=rank * N_loc)
    yv = torch.empty_like(psi)
    nrm2 = vec_dot(psi, psi).real
    if world > 1:
        t = torch.tensor([nrm2], dtype=torch.float64, device='cuda')
        dist.all_reduce(t)
        nrm2 = t.item()
        if info['n_global_parts'] > 0:
            free_b, _ = torch.cuda.mem_get_info()
            # two exchange buffers: more were measured to bring nothing (the exchanges already hide
            # behind the local part) and only raise the memory footprint
            nbuf = 2 if free_b > 3.2 * 16 * N_loc else 1
            recv = torch.empty(nbuf * N_loc, dtype=torch.complex128, device='cuda')
    vec_scale(psi, 1.0 / np.sqrt(nrm2))

    def step():
        if world > 1:
            return eng.apply_sharded(psi, out=yv, recv=recv, want_expect=True)[1]
        return eng.apply(psi, out=yv, want_expect=True)[1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    energy = None
    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        energy = step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for _ in range(args.steps):
            energy = step()
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
    ms_step = ms_total / args.steps
    value = T * float(1 << n_total) / (ms_step * 1e-3) / 1e9

    # ---- e2e: public API, host buffers -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from qiskit_aqua_b200.operators import Pauli
        if world == 1:
            op = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(xx), int(zz), n_total)]
                                               for xx, zz, cc in zip(x, z, c)])
            op._engine = eng                       # same packed operator; packing is a one-off
        try:
            host = torch.empty(N_loc, dtype=torch.complex128, pin_memory=True)
        except RuntimeError as ex:                       # host cannot pin 16 GiB per rank
            host = None
            e2e = {'unavailable': 'pinned host allocation failed: %s' % str(ex)[:80]}
        if host is not None:
            host.copy_(psi)
            torch.cuda.synchronize()

            def e2e_step():
                # the shim copies the pinned host state to the device (H2D) and reads the scalar back
                if world == 1:
                    return op.evaluate_with_statevector(host)[0]
                # y is not stored in this mode: its buffer receives the shard; the copy is pipelined
                # with the chunk-local passes (psum_expect_sharded_host)
                return eng.expect_sharded_host(host, scratch=yv, recv=recv)

            e2e_step()
            barrier()
            ksteps = max(2, min(args.steps, 3))
            ev0.record()
            for _ in range(ksteps):
                e_val = e2e_step()
            ev1.record()
            barrier()
            ms_e = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms_e], dtype=torch.float64, device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_e = t.item()
            ms_e /= ksteps
            e2e = {'value': T * float(1 << n_total) / (ms_e * 1e-3) / 1e9, 'unit': UNIT,
                   'h2d_bytes_per_step': 16 * N_loc * world, 'd2h_bytes_per_step': 16 * world,
                   'ms_per_step': ms_e, 'api': 'WeightedPauliOperator.evaluate_with_statevector'
                   if world == 1 else 'PauliSumEngine.expect_sharded_host',
                   'energy': e_val.real}

    # ---- roofline -------------------------------------------------------------------------------
    peak, peak_src = peaks()
    b_alg = 16.0 * N_loc * (gx_total + 1)
    achieved = b_alg / (ms_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath) and world == 1 and args.workload == 'c4' and n_loc == 30:
        try:
            traffic = json.load(open(tpath)).get('dram_bytes_per_step')
        except Exception:
            traffic = None
    # secondary, for context: what the DRAM and the shared-memory pipe actually see
    dram_gbs = (traffic / (ms_step * 1e-3) / 1e9) if traffic else None
    sm_mhz = 1965.0
    smem_floor_ms = info['n_group_records'] * float(N_loc) * 16.0 / (148 * 128.0 * sm_mhz * 1e6) * 1e3
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'dram_achieved': dram_gbs, 'dram_frac': (dram_gbs / peak) if dram_gbs else None,
                'smem_floor_ms': smem_floor_ms, 'smem_floor_frac': smem_floor_ms / ms_step,
                'traffic': traffic, 'peak_source': peak_src, 'kernel': 'psum::k_pass (x%d launches per step)' % info['n_passes'],
                'algorithmic_bytes_per_step': b_alg,
                'note': 'B_alg = 16*N_loc*(G_x+1) (SURVEY 8d) counts one partner read per x-group; k_pass serves '
                        'in-tile groups from shared memory, so DRAM traffic is far below B_alg and frac exceeds 1; '
                        'dram_* = ncu DRAM bytes / time, smem_floor = 16 B per (group record, amplitude) at 128 B/clk/SM'}
    launches = (info['n_passes'] + 1) * args.steps

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': ('C4 ising+heisenberg+3local' if args.workload == 'c4' else 'C5 random weight<=4 paulis') +
                       ' %dq %d terms, apply+expect' % (n_total, T),
                       'qubits': n_total, 'terms': T, 'x_groups': gx_total, 'passes': info['n_passes'],
                       'patterns': info['n_patterns'], 'remote_groups': info['n_remote_groups'],
                       'tile_bits': info['tile_bits'], 'sharding': 'top %d qubits' % p_bits,
                       'global_parts': info['n_global_parts'],
                       'l2': 'inputs (%.0f GiB per GPU) larger than L2' % (16 * N_loc / 2 ** 30)},
            'energy': energy.real if energy is not None else None,
            'clocks': clk.summary(), 'gpu_launches': launches, 'roofline': roofline}
    if e2e:
        line['e2e'] = e2e

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        psi_host = (host if (not args.no_e2e and host is not None) else psi.cpu()).numpy()
        line['cpu_baseline'] = cpu_sample(n_total, x, z, y, c, psi_host)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
