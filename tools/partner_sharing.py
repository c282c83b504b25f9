"""Planning study (CPU only): how many partner loads could be shared if x-groups that differ only in
the register bits of a pass reused one set of 8 partner amplitudes ("load sets" vs groups).
See DESIGN.md, candidates for the next round."""
import sys, itertools; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qiskit_aqua_b200.engine import PauliSumEngine
from qiskit_aqua_b200 import workloads as W
def analyse(name, n, x, z, y, c, nreg=3):
    e = PauliSumEngine(n, x, z, y, c)
    info = e.info()
    xs = sorted(set(int(v) for v in x))
    tot_g = tot_sets = 0
    covered=set()
    for p in range(info['n_passes']):
        fm = e.pass_info(p)['free_mask']
        members = [m for m in xs if (m & ~fm) == 0 and m not in covered]
        covered.update(members)
        fb = [b for b in range(n) if fm >> b & 1][3:]   # three lowest are lane bits
        best = None
        for R in itertools.combinations(fb, nreg):
            rm = sum(1 << b for b in R)
            sets = len(set(m & ~rm for m in members if m & ~rm))
            if best is None or sets < best: best = sets
        tot_g += len(members); tot_sets += best
    print(name, 'groups', tot_g, 'load sets', tot_sets, 'ratio %.2f' % (tot_sets / max(1,tot_g)), 'passes', info['n_passes'], 'remote', info['n_remote_groups'])
_, x, z, y, c = bench.workload(30, 'c4'); analyse('c4-30q', 30, x, z, y, c); analyse('c4-30q/4reg', 30, x, z, y, c, 4)
_, x, z, y, c = bench.workload(31, 'c5'); analyse('c5-31q', 31, x, z, y, c)
for nq in (14, 20):
    r = W.molecular_style(nq); analyse('molecular-%dq' % nq, nq, *r[1:])
