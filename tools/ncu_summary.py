#!/usr/bin/env python
"""Turns ncu CSV exports (gpurun_out/) into the small tracked summaries under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches_r1_q30.csv profiles/launches_r1_q30.md [steps_kernels]
  python tools/ncu_summary.py raw gpurun_out/raw.csv profiles/ncu_full_r1_q26.md
"""
import collections
import csv
import json
import os
import sys


def launches(path, out, per_step):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    ui = hdr.index('Metric Unit')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[ii]), {'kernel': r[ki].split('(')[0]})
        d[r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    seq = list(per.values())
    kp = [d for d in seq if 'k_pass' in d['kernel']]
    step = kp[-per_step:]
    tot_t = sum(d['gpu__time_duration.sum'] for d in seq)
    lines = ['# ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum '
             '--clock-control none`)', '', 'source: `%s` (cold-cache, serialised; compare SHARES not absolutes)' % path, '',
             '| kernel | launches | total ms | share |', '|---|---|---|---|']
    agg = collections.OrderedDict()
    for d in seq:
        a = agg.setdefault(d['kernel'], [0, 0.0])
        a[0] += 1
        a[1] += d['gpu__time_duration.sum']
    for k, (c, t) in agg.items():
        lines.append('| `%s` | %d | %.3f | %.2f%% |' % (k, c, t / 1e6, 100 * t / tot_t))
    lines += ['', '## last H.psi step: the %d `k_pass` launches (one per pass)' % per_step, '',
              '| pass | ms | DRAM read GB | DRAM write GB | DRAM GB/s |', '|---|---|---|---|---|']
    tb = 0.0
    for i, d in enumerate(step):
        rd, wr, t = d.get('dram__bytes_read.sum', 0.0), d.get('dram__bytes_write.sum', 0.0), d['gpu__time_duration.sum']
        unit = 1e-9
        tb += rd + wr
        lines.append('| %d | %.2f | %.2f | %.2f | %.0f |' % (i, t / 1e6, rd * unit, wr * unit, (rd + wr) * unit / (t / 1e9) if t else 0))
    tstep = sum(d['gpu__time_duration.sum'] for d in step) / 1e6
    lines += ['', 'step total: %.1f ms under ncu, DRAM traffic %.1f GB per step' % (tstep, tb / 1e9)]
    open(out, 'w').write('\n'.join(lines) + '\n')
    return {'dram_bytes_per_step': tb, 'ms_per_step_under_ncu': tstep,
            'per_launch_dram_bytes': [d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in step]}


WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def raw(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = ['# ncu --set full summary (`--page raw`)', '', 'source: `%s`; one column per `k_pass` launch (pass)' % path, '',
             '| metric | unit | ' + ' | '.join('p%d' % i for i in range(len(data))) + ' |',
             '|---|---|' + '---|' * len(data)]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals = []
            for r in data:
                try:
                    vals.append('%.3g' % float(r[i].replace(',', '')))
                except ValueError:
                    vals.append(r[i])
            lines.append('| `%s` | %s | %s |' % (w, units[i], ' | '.join(vals)))
    open(out, 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        info = launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 17)
        tj = os.path.join(os.path.dirname(sys.argv[3]), 'traffic.json')
        info['source'] = sys.argv[3]
        json.dump(info, open(tj, 'w'), indent=1)
        print(json.dumps(info)[:300])
    else:
        raw(sys.argv[2], sys.argv[3])
