"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, count, share)."""
import collections
import csv
import io
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.DictReader(io.StringIO("".join(lines)))
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'n':>6s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k[:60]:60s} {cnt[k]:6d} {v:11.1f} {v / cnt[k]:9.2f} {v / T * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
