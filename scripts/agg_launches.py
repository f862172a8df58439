"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("%-72s %6s %12s %6s" % ("kernel", "count", "total_us", "share"))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%-72s %6d %12.1f %5.1f%%" % (k, c, t, 100 * t / tot))
print("%-72s %6d %12.1f" % ("TOTAL", sum(c for c, _ in agg.values()), tot))
