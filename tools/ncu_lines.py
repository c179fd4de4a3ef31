"""Join an `ncu --page source --csv` (SASS view) export with `nvdisasm -g` line info: stall samples per CUDA source line.

    python tools/ncu_lines.py <sass_page.csv> <nvdisasm -g output> <mangled kernel name substring> [top_n]
"""
import collections
import csv
import re
import sys

page, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 45
rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows[:5]) if "# Samples" in r)
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ix["Address"]], 16)
# offset -> (file, line) from the disassembly of that kernel
loc, cur, inside = {}, ("?", 0), False
for ln in open(dis):
    if ln.startswith(".text."):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        loc[int(m.group(1), 16)] = cur
tot = 0
per = collections.defaultdict(lambda: collections.Counter())
for r in data:
    off = int(r[ix["Address"]], 16) - base
    f = loc.get(off, ("?", 0))
    n = int(r[ix["# Samples"]] or 0)
    tot += n
    per[f]["samples"] += n
    per[f]["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for s in stalls:
        per[f][s] += int(r[ix[s]] or 0)
print("total samples", tot)
agg = collections.Counter()
for f, c in per.items():
    for s in stalls:
        agg[s] += c[s]
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in agg.most_common(8)))
byfile = collections.Counter()
for f, c in per.items():
    byfile[f[0]] += c["samples"]
print("by file:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byfile.most_common(8)))
for f, c in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
    st = sorted(((c[s], s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{f[0]:18s}:{f[1]:4d} {100 * c['samples'] / tot:5.1f}% inst {c['inst']:9d}  " +
          " ".join(f"{n}:{100 * v / max(1, c['samples']):.0f}%" for v, n in st))
