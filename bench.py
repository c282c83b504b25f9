#!/usr/bin/env python
"""bench.py -- headline benchmark of the Pauli-sum H.psi / <psi|H|psi> hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): Pauli-sum H.psi throughput in G term-amplitudes/s on the 30-qubit,
10 000-term transverse-field Ising + Heisenberg (+3-local) Hamiltonian (config C4).  One "step" =
one y = H.psi with the fused <psi|H|psi> reduction over a statevector resident in HBM.  With N > 1
ranks (torchrun) the statevector has 30 + log2(N) qubits, sharded by its top qubits (weak scaling:
2^30 amplitudes per GPU), with the NCCL amplitude exchange inside the step.

The JSON line also carries
  parity       measured OUTSIDE the timed region on the very buffers the timed steps produced:
               256 sampled rows of the (sharded) y against the CPU oracle
               (oracle.c_oracle.apply_rows_counter evaluates rows of H.psi of the counter-based
               state without a 2^n host array), the fused <psi|H|psi> against an independent
               reduction of <psi|y>, and <H> of a seeded PRODUCT state against its closed form
               (oracle.pauli_oracle.product_state_expect, O(T n)) -- a full-size energy check
  e2e          the same metric through the public API with a pageable numpy state
               (WeightedPauliOperator.evaluate_with_statevector(ndarray): the reference's own call);
               the pinned-host variant sits beside it.  H2D of psi and D2H of the scalar are inside
  roofline     frac = max(DRAM floor, shared-memory-pipe floor) / measured time, both floors computed
               from the plan the kernels execute (and validated against the ncu capture under
               profiles/); B_alg = 16 N (G_x + 1) of SURVEY 8d is kept as a secondary field
  nvlink       (N > 1) bytes exchanged per rank / time the exchanges occupy the communication
               stream vs 900 GB/s per direction, and how much of that time hides behind the passes
  cpu_baseline the oracle C restatement (oracle/psum_oracle.c, OpenMP) on a bounded row sample of
               the same workload, plus the literal reference semantics (one CSR per Pauli -> sum ->
               csr.dot) at a size it can still assemble
"""
import argparse
import ctypes as C
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
H2D_GBS = 55.6      # pinned H2D copy rate measured on this pool's B200 hosts (tools/h2d_probe.py, profiles/)
NVLINK_GBS = 900.0          # NVLink 5, per direction per GPU
SM_CLOCK_MHZ = 1965.0
NUM_SMS = 148


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--qubits', type=int, default=0, help='override qubits per GPU (default 30)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--workload', default='c4', choices=['c4', 'c5'],
                    help='c4: Ising+Heisenberg+3-local (headline); c5: random weight<=4 Pauli sum')
    ap.add_argument('--tile-bits', type=int, default=0)
    ap.add_argument('--low-bits', type=int, default=0)
    ap.add_argument('--reg-bits', type=int, default=0)
    ap.add_argument('--cta-mode', type=int, default=-1)
    ap.add_argument('--late-split', type=int, default=-1)
    ap.add_argument('--host-chunk-bits', type=int, default=0)
    ap.add_argument('--copy-threads', type=int, default=0)
    ap.add_argument('--min-bundle', type=int, default=0)
    ap.add_argument('--plan-tries', type=int, default=0)
    ap.add_argument('--term-cache', type=int, default=-1)
    ap.add_argument('--min-cover', type=int, default=0)
    ap.add_argument('--share-weight', type=int, default=-1)
    ap.add_argument('--comm-sms', type=int, default=-1)
    ap.add_argument('--peer-copy', type=int, default=-1)
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


# ------------------------------------------------------------------------------------------------
# CPU legs (test infrastructure: the oracle is only ever the checker / the timed baseline)
# ------------------------------------------------------------------------------------------------
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


def literal_reference_leg(n=16, n_terms=300, budget_s=20.0):
    """The reference's literal arithmetic (legacy/op_converter.py:120-123: one CSR per Pauli, summed;
    legacy/weighted_pauli_operator.py:621-622: csr.dot + vdot) on the first `n_terms` terms of the
    C4-style operator at n qubits -- a size it can still assemble.  Build and matvec timed separately;
    numpy/scipy are single-threaded here."""
    from oracle import pauli_oracle as po
    from qiskit_aqua_b200 import workloads as W
    _, x, z, yp, c = W.ising_heisenberg(n, n_triples=max(0, n_terms), seed=30)
    x, z, yp, c = x[:n_terms], z[:n_terms], yp[:n_terms], c[:n_terms]
    terms = [(complex(cc), po.masks_to_label(int(xx), int(zz), n)) for xx, zz, cc in zip(x, z, c)]
    t0 = time.perf_counter()
    mat = po.ref_to_spmatrix(terms)
    t_build = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    reps, t_mv = 0, 0.0
    e = 0.0
    while reps < 3 or (t_mv < 1.0 and t_build + t_mv < budget_s):
        t0 = time.perf_counter()
        e = np.vdot(psi, mat.dot(psi))
        t_mv += time.perf_counter() - t0
        reps += 1
    t_mv /= reps
    T = len(terms)
    return {'qubits': n, 'terms': T, 'nnz': int(mat.nnz), 'build_s': t_build, 'matvec_ms': 1e3 * t_mv,
            'value_matvec_only': T * float(1 << n) / t_mv / 1e9,
            'value_as_the_reference_runs_it': T * float(1 << n) / (t_build + t_mv) / 1e9,
            'unit': UNIT, 'cores': 1, 'energy': float(np.real(e)),
            'note': 'evaluate_with_statevector rebuilds the CSR on every call (SURVEY 8 a2/a3): the second '
                    'value is what a caller of the reference sees'}


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
              'the host), oracle/psum_oracle.c OpenMP on %d threads' % (rows, T, n_total, n_cpu, co.threads()))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C4 ising+heisenberg+3local %dq %d terms (row sample)' % (n_total, T),
                       'qubits': n_total, 'terms': T},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': co.threads(), 'kind': 'port',
                             'sample': sample,
                             'note': 'the ratio to the GPU arm scales with this host\'s core count; the literal '
                                     'CSR path of the reference is reported by the GPU arm under '
                                     'cpu_baseline.literal_reference'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# floors computed from the plan
# ------------------------------------------------------------------------------------------------
def plan_floors(eng, hbm_gbs, store=True):
    """DRAM bytes and shared-memory wavefronts of one H.psi on this rank (psum_plan_floors walks the
    plan records the kernels execute) and the two time floors they imply."""
    f = eng.plan_floors(store)
    t_dram = f['dram_bytes'] / (hbm_gbs * 1e9)
    t_smem = f['smem_wavefronts'] / (NUM_SMS * SM_CLOCK_MHZ * 1e6)
    return f['dram_bytes'], f['smem_wavefronts'], t_dram, t_smem


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
    if args.cta_mode >= 0:
        opts['cta_mode'] = args.cta_mode
    if args.min_bundle:
        opts['min_bundle'] = args.min_bundle
    if args.plan_tries:
        opts['plan_tries'] = args.plan_tries
    if args.term_cache >= 0:
        opts['term_cache'] = args.term_cache
    if args.min_cover:
        opts['min_cover'] = args.min_cover
    if args.share_weight >= 0:
        opts['share_weight'] = args.share_weight
    eng = PauliSumEngine(n_total, x, z, y, c, opts)
    recv = None
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.shard(rank, world, uid[0])
    if args.comm_sms >= 0:
        eng.set_option('comm_sms', args.comm_sms)
    if args.peer_copy >= 0:
        eng.set_option('peer_copy', args.peer_copy)
    if args.late_split >= 0:
        eng.set_option('late_split', args.late_split)
    if args.host_chunk_bits:
        eng.set_option('host_chunk_bits', args.host_chunk_bits)
    if args.copy_threads:
        eng.set_option('copy_threads', args.copy_threads)
    info = eng.info()
    gx_total = info['n_xgroups']

    def allreduce(v, op=None):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return t.item()

    SEED = 30
    psi = fill_state(N_loc, seed=SEED, first_index=rank * N_loc)
    yv = torch.empty_like(psi)
    nrm2 = allreduce(vec_dot(psi, psi).real)
    if world > 1 and info['n_global_parts'] > 0:
        free_b, _ = torch.cuda.mem_get_info()
        # one receive buffer per exchanged shard when memory allows (the partner shards are pulled by
        # the copy engines while the local part runs, so every pull can complete behind it); the ring
        # works with any count >= 1.  One shard of head-room stays free for the e2e scratch.
        nbuf = int(max(1, min(info['n_global_parts'], (free_b - 1.5 * 16 * N_loc) // (16 * N_loc))))
        recv = torch.empty(nbuf * N_loc, dtype=torch.complex128, device='cuda')
    scale = 1.0 / np.sqrt(nrm2)
    vec_scale(psi, scale)

    def step():
        if world > 1:
            return eng.apply_sharded(psi, out=yv, recv=recv, want_expect=True)[1]
        return eng.apply(psi, out=yv, want_expect=True)[1]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        """nsteps calls of fn between barriers, CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(nsteps):
            out = fn()
        e1.record()
        barrier()
        return allreduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None) / nsteps, out

    energy = None
    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        energy = step()
    with ClockSampler(local_rank) as clk:
        ms_step, energy = timed(step, args.steps)
    value = T * float(1 << n_total) / (ms_step * 1e-3) / 1e9

    # ---- parity (outside the timed region, on the buffers the timed steps left behind) ----------
    parity = None
    if not args.no_parity:
        from oracle import c_oracle as co
        from oracle import pauli_oracle as po
        co.use_all_cores()
        rng = np.random.default_rng(12345)
        rows = np.unique(rng.integers(0, 1 << n_total, size=256, dtype=np.uint64))
        mine = rows[(rows >> np.uint64(n_loc)) == np.uint64(rank)]
        err_y, ref_max = 0.0, 0.0
        if len(mine):
            got = yv[torch.from_numpy((mine & np.uint64(N_loc - 1)).astype(np.int64)).cuda()].cpu().numpy()
            ref = co.apply_rows_counter(x, z, y, c, SEED, scale, mine)
            err_y, ref_max = float(np.abs(got - ref).max()), float(np.abs(ref).max())
        ref_max = allreduce(ref_max, dist.ReduceOp.MAX if world > 1 else None)
        err_y = allreduce(err_y, dist.ReduceOp.MAX if world > 1 else None) / max(ref_max, 1e-300)
        # the fused reduction against an independent reduction of <psi|y>
        e_dot = allreduce(vec_dot(psi, yv).real)
        err_fused = abs(energy.real - e_dot) / max(abs(e_dot), 1e-300)
        parity = {'rows': int(len(rows)), 'max_rel_err_y': err_y, 'rel_err_energy_fused_vs_dot': err_fused,
                  'oracle': 'oracle/psum_oracle.c apply_rows_counter (rows of H.psi of the seeded counter state)'}

    # ---- e2e: public API, host buffers -----------------------------------------------------------
    e2e = None
    host_np = None
    if not args.no_e2e:
        from qiskit_aqua_b200.operators import Pauli
        op = None
        if world == 1:
            op = WeightedPauliOperator(paulis=[[cc, Pauli.from_masks(int(xx), int(zz), n_total)]
                                               for xx, zz, cc in zip(x, z, c)])
            op._engine = eng                       # same packed operator; packing is a one-off
        host_np = np.empty(N_loc, dtype=np.complex128)          # pageable: what the reference API receives
        host_np[:] = psi.cpu().numpy()
        ksteps = max(2, args.steps)

        def e2e_pageable():
            if world == 1:
                return op.evaluate_with_statevector(host_np)[0]
            return eng.expect_sharded_host(host_np, scratch=yv, recv=recv)

        e2e_pageable()
        ms_pg, e_pg = timed(e2e_pageable, ksteps)
        e2e = {'value': T * float(1 << n_total) / (ms_pg * 1e-3) / 1e9, 'unit': UNIT,
               'h2d_bytes_per_step': 16 * N_loc * world, 'd2h_bytes_per_step': 16 * world,
               'ms_per_step': ms_pg, 'steps': ksteps, 'host_memory': 'pageable numpy.ndarray',
               'api': 'WeightedPauliOperator.evaluate_with_statevector(numpy.ndarray)'
               if world == 1 else 'PauliSumEngine.expect_sharded_host(numpy.ndarray)',
               'host_gbs': 16.0 * N_loc * world / (ms_pg * 1e-3) / 1e9,
               'energy': e_pg.real,
               'bound': 'host memory / PCIe: every rank of the node pulls its 16 GiB shard through the same '
                        'host memory system (aggregate host_gbs is the measured ceiling)'}
        try:
            host_pin = torch.empty(N_loc, dtype=torch.complex128, pin_memory=True)
        except RuntimeError as ex:                       # host cannot pin 16 GiB per rank
            host_pin = None
            e2e['pinned'] = {'unavailable': 'pinned host allocation failed: %s' % str(ex)[:80]}
        if host_pin is not None:
            host_pin.copy_(torch.from_numpy(host_np))

            def e2e_pinned():
                if world == 1:
                    return op.evaluate_with_statevector(host_pin)[0]
                return eng.expect_sharded_host(host_pin, scratch=yv, recv=recv)

            e2e_pinned()
            ms_pin, e_pin = timed(e2e_pinned, ksteps)
            e2e['pinned'] = {'value': T * float(1 << n_total) / (ms_pin * 1e-3) / 1e9, 'ms_per_step': ms_pin,
                             'host_gbs': 16.0 * N_loc * world / (ms_pin * 1e-3) / 1e9, 'energy': e_pin.real}
            e2e['pageable_over_pinned'] = ms_pg / ms_pin
            del host_pin
        if world == 1:      # the same expectation with psi already resident: what the copy has to hide behind
            eng.expect(psi)
            ms_dev, _ = timed(lambda: eng.expect(psi), ksteps)
            e2e['device_expect_ms'] = ms_dev
            e2e['h2d_floor_ms'] = 16.0 * N_loc / (H2D_GBS * 1e9) * 1e3
            e2e['h2d_peak_gbs'] = H2D_GBS
        if parity is not None:
            parity['rel_err_energy_e2e_vs_device'] = abs(e_pg.real - energy.real) / max(abs(energy.real), 1e-300)

    # ---- full-size energy parity: <H> of a seeded product state vs its closed form ----------------
    if parity is not None:
        from oracle import pauli_oracle as po
        from qiskit_aqua_b200.circuits import StatevectorCircuit
        rngp = np.random.default_rng(777)
        th, ph = rngp.uniform(0.2, np.pi - 0.2, n_total), rngp.uniform(0, 2 * np.pi, n_total)
        a = np.exp(-0.5j * ph) * np.cos(th / 2)
        b = np.exp(0.5j * ph) * np.sin(th / 2)
        circ = StatevectorCircuit(n_loc)
        for q in range(n_loc):
            circ.ry(float(th[q]), q)
            circ.rz(float(ph[q]), q)
        circ.run(out=psi)                                  # psi is no longer needed: reuse its buffer
        fac = 1.0 + 0.0j
        for q in range(n_loc, n_total):
            fac *= b[q] if (rank >> (q - n_loc)) & 1 else a[q]
        vec_scale(psi, fac)
        e_prod_apply = step()
        if world > 1:
            e_prod_expect = eng.apply_sharded(psi, recv=recv, want_expect=True, store=False)[1]
        else:
            e_prod_expect = eng.expect(psi)
        e_ref = po.product_state_expect(n_total, x, z, y, c, a, b)
        den = max(abs(e_ref), 1e-300)
        parity['rel_err_energy_product_state'] = max(abs(e_prod_apply - e_ref), abs(e_prod_expect - e_ref)) / den
        parity['energy_product_state'] = e_ref.real
        parity['product_state_oracle'] = 'oracle/pauli_oracle.py product_state_expect (closed form, O(T n))'
        parity['ok'] = bool(parity['max_rel_err_y'] <= 1e-10 and parity['rel_err_energy_product_state'] <= 1e-10 and
                            parity['rel_err_energy_fused_vs_dot'] <= 1e-10)

    # ---- sharded Lanczos == single-GPU Lanczos (the on-device eigensolver on the sharded matvec) ----
    if parity is not None and world > 1:
        from qiskit_aqua_b200 import workloads as W
        nl, xl_, zl_, yl_, cl_ = W.two_local_random(22, 700, seed=22)
        engl = PauliSumEngine(nl, xl_, zl_, yl_, cl_)
        uid2 = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        engl.shard(rank, world, uid2[0])
        rs = engl.lanczos(k=2, tol=1e-12, want_residuals=True)
        single = [0.0, 0.0]
        if rank == 0:
            single = PauliSumEngine(nl, xl_, zl_, yl_, cl_).lanczos(k=2, tol=1e-12)['evals'].tolist()
        t = torch.tensor(single, dtype=torch.float64, device='cuda')
        dist.broadcast(t, src=0)
        single = t.cpu().numpy()
        parity['sharded_lanczos'] = {'qubits': nl, 'terms': len(xl_), 'k': 2, 'tol': 1e-12, 'matvecs': rs['iters'],
                                     'evals': rs['evals'].tolist(),
                                     'max_abs_err_vs_single_gpu': float(np.abs(rs['evals'] - single).max()),
                                     'max_residual': float(rs['resid'].max())}
        parity['ok'] = bool(parity['ok'] and parity['sharded_lanczos']['max_abs_err_vs_single_gpu'] <= 1e-9 and
                            parity['sharded_lanczos']['max_residual'] <= 1e-6)
        del engl

    # ---- NVLink: exchange time on the communication stream, and how much of it is hidden ----------
    nvlink = None
    if world > 1 and info['n_global_parts'] > 0:
        eng.set_option('profile_exchange', 1)
        step()
        stats = eng.exchange_stats()
        eng.set_option('profile_exchange', 0)
        eng.set_option('skip_exchange', 1)          # same launches, no NCCL traffic: compute only
        step()
        ms_noex, _ = timed(step, max(2, min(args.steps, 3)))
        eng.set_option('skip_exchange', 0)
        ex_ms = allreduce(stats[2], dist.ReduceOp.MAX)
        ex_bytes = stats[0] * stats[1]
        exposed = max(0.0, ms_step - ms_noex)
        nvlink = {'exchanges_per_step': int(stats[0]), 'bytes_per_rank_per_direction': ex_bytes,
                  'exchange_ms_on_comm_stream': ex_ms,
                  'achieved_gbs_per_direction': ex_bytes / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None,
                  'peak_gbs_per_direction': NVLINK_GBS,
                  'frac': (ex_bytes / (ex_ms * 1e-3) / 1e9 / NVLINK_GBS) if ex_ms > 0 else None,
                  'step_ms': ms_step, 'step_ms_without_exchange': ms_noex, 'exposed_ms': exposed,
                  'overlap_frac': (1.0 - exposed / ex_ms) if ex_ms > 0 else None,
                  'note': 'exchange_ms = sum of per-exchange CUDA-event intervals on the communication stream '
                          '(includes waiting for the peer and for a free receive slot); overlap_frac = share of '
                          'it hidden behind the passes (step time with vs without the NCCL calls)'}

    # ---- roofline ---------------------------------------------------------------------------------
    peak, peak_src = peaks()
    b_alg = 16.0 * N_loc * (gx_total + 1)
    try:
        dram_b, wf, t_dram, t_smem = plan_floors(eng, peak)
    except Exception as ex:                                           # never lose the line over the model
        dram_b, wf, t_dram, t_smem = None, None, 0.0, 0.0
        sys.stderr.write('plan_floors failed: %r\n' % (ex,))
    t_meas = ms_step * 1e-3
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath) and world == 1:
        try:
            tj = json.load(open(tpath))
            if tj.get('workload') == args.workload and tj.get('qubits') == n_loc and tj.get('passes') == info['n_passes'] \
                    and tj.get('reg_bits', 3) == (args.reg_bits or 3):
                traffic = tj.get('dram_bytes_per_step')
        except Exception:
            traffic = None
    bound = 'hbm' if t_dram >= t_smem else 'smem-pipe'
    floor = max(t_dram, t_smem)
    roofline = {'bound': bound, 'frac': floor / t_meas if t_meas > 0 else None,
                'achieved': (dram_b / t_meas / 1e9) if dram_b else None, 'peak': peak, 'unit': 'GB/s',
                'dram_frac': (dram_b / t_meas / 1e9 / peak) if dram_b else None,
                'traffic': traffic, 'traffic_model': dram_b,
                'dram_floor_ms': 1e3 * t_dram, 'smem_floor_ms': 1e3 * t_smem,
                'smem_wavefronts_per_step': wf,
                'smem_pipe_frac': (1e3 * t_smem / ms_step) if ms_step > 0 else None,
                'peak_source': peak_src,
                'kernel': 'psum::k_pass (x%d launches per step)' % info['n_passes'],
                'algorithmic_bytes_per_step': b_alg, 'b_alg_gbs': b_alg / t_meas / 1e9,
                'note': 'frac = max(DRAM floor, shared-memory-pipe floor) / measured step; achieved = DRAM bytes the '
                        'plan moves (validated by ncu, profiles/) / time; the kernel serves in-tile partners from '
                        'shared memory, so SURVEY 8d\'s B_alg = 16 N (G_x+1) (b_alg_gbs) is not a DRAM figure'}
    launches = (info['n_passes'] + 1) * args.steps

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': ('C4 ising+heisenberg+3local' if args.workload == 'c4' else 'C5 random weight<=4 paulis') +
                       ' %dq %d terms, apply+expect' % (n_total, T),
                       'qubits': n_total, 'terms': T, 'x_groups': gx_total, 'passes': info['n_passes'],
                       'group_records': info['n_group_records'], 'partner_load_sets': info['n_bundles'],
                       'patterns': info['n_patterns'], 'remote_groups': info['n_remote_groups'],
                       'tile_bits': info['tile_bits'], 'reg_bits': args.reg_bits or 3,
                       'sharding': 'top %d qubits' % p_bits, 'global_parts': info['n_global_parts'],
                       'recv_buffers': (recv.numel() // N_loc) if recv is not None else 0,
                       'l2': 'inputs (%.0f GiB per GPU) larger than L2' % (16 * N_loc / 2 ** 30)},
            'energy': energy.real if energy is not None else None,
            'clocks': clk.summary(), 'gpu_launches': launches, 'roofline': roofline}
    if parity is not None:
        line['parity'] = parity
    if nvlink is not None:
        line['nvlink'] = nvlink
    if e2e:
        line['e2e'] = e2e

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if host_np is None:
            from oracle import c_oracle as co
            host_np = co.counter_state(SEED, 0, N_loc) * scale
        line['cpu_baseline'] = cpu_sample(n_total, x, z, y, c, host_np)
        try:
            line['cpu_baseline']['literal_reference'] = literal_reference_leg()
        except Exception as ex:
            line['cpu_baseline']['literal_reference'] = {'unavailable': repr(ex)[:120]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
