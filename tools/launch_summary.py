"""Sums an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel: launches, total and mean duration,
share of the summed GPU time.  usage: launch_summary.py <launches.csv>"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void\s+", "", name)
    rows.append((name, v))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1; agg[n][1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms summed")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / tot:6.1%} {t / 1e3:10.2f} ms {c:8d} x {t / c:9.1f} us  {n[:110]}")
