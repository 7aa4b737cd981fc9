"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel (development aid).
usage: python tools/launch_list.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ik][:100], []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-102s n=%4d total %9.3f ms %5.1f%%  last: %s" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot, " ".join("%.2f" % (x / 1e6) for x in v[-4:])))
print("total %.3f ms" % (tot / 1e6))
