"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); u = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(u, 1)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), ns))
tot = sum(ns for _, ns in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    agg[k][0] += 1; agg[k][1] += ns
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total (ncu-serialised, cold cache: compare SHARES)")
print(f"{'kernel':48s} {'count':>6s} {'ms':>10s} {'share':>7s}")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:48]:48s} {c:6d} {ns / 1e6:10.3f} {100 * ns / tot:6.1f}%")
