"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total / count / share."""
import csv, sys, collections, re
def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(u, 1.0)
        name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""); name = re.sub(r"\(.*", "", name); name = re.sub(r"<.*", "", name)
        rows.append((name, v))
    return rows
for p in sys.argv[1:]:
    rows = load(p)
    agg = collections.OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0.0, 0]); a[0] += v; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print(f"## {p}: {len(rows)} launches, {tot:.1f} us total")
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t:10.1f} us  {100*t/tot:5.1f}%  x{c:<3d} avg {t/c:9.1f} us  {n[:90]}")
