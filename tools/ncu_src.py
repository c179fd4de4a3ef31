"""Summarise an `ncu --page source --csv` export: stall mix and the hottest SASS instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {s: sum(num(r, s) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f'{s:28s}{v:8d} {100 * v / tot:5.1f}%')
for pos, r in enumerate(data):
    r.append(pos)
top = sorted(data, key=lambda r: -num(r, '# Samples'))[:top_n]
for r in top:
    st = sorted(((num(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"#{r[-1]:5d}", f"{num(r, '# Samples'):6d}", r[ix['Instructions Executed']].rjust(8), r[ix['Source']][:64].ljust(64), st)
