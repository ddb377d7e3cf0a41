"""Summarise an `ncu --page source --csv` export: top instructions by warp-stall samples.
usage: python profiles/ncu_stalls.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
i_src, i_n = hdr.index('Source'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_')]
data, tot = [], 0
for r in rows[2:]:
    try:
        s = int(r[i_n])
    except (ValueError, IndexError):
        continue
    tot += s
    data.append((s, r))
print('kernel:', rows[0][1][:100])
print('total samples', tot, 'instructions', len(data))
agg = {}
for s, r in data:
    for i, h in stall_cols:
        if r[i].isdigit():
            agg[h] = agg.get(h, 0) + int(r[i])
print('by reason:', sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True)[:8])
data.sort(key=lambda x: -x[0])
for s, r in data[:top]:
    reasons = sorted([(int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols], reverse=True)[:2]
    print(f"{s:5d} {100 * s / max(tot, 1):5.1f}%  {r[i_src][:78]:78s} {[(h[6:], n) for n, h in reasons if n > 0]}")
