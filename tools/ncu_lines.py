"""Rank CUDA source lines of one launch by warp-stall samples.
usage: ncu -i rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > f.csv; python tools/ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
hdr = rows[hi]
ci = {}
for i, h in enumerate(hdr):
    ci.setdefault(h, i)
S, I = ci['# Samples'], ci['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) <= S or r[2] != '-':      # per-line aggregate rows have '-' as address
        continue
    try:
        data.append((int(r[S]), r))
    except ValueError:
        pass
tot = sum(s for s, _ in data)
print('total samples', tot)
agg = {}
for s, r in data:
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(r[ci[h]] or 0)
print('stall mix:', ' '.join('%s:%.1f%%' % (h[6:], 100.0 * v / max(1, sum(agg.values()))) for h, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
for s, r in sorted(data, key=lambda t: -t[0])[:top_n]:
    top = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print('%5s %5.1f%% inst=%-11s %-44s | %s' % (r[0], 100.0 * s / tot, r[I], ' '.join('%s:%d' % (h, v) for v, h in top), r[1].strip()[:100]))
