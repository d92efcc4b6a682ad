#!/usr/bin/env python3
"""ncu launch list (gpu__time_duration.sum, --csv) -> total ms per kernel name over the LAST MSM in the capture (tools)."""
import collections, csv, re, sys
rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
ls = []
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    ls.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ab200::", ""), ms, r["Grid Size"]))
starts = [i for i, l in enumerate(ls) if l[0].startswith("msm_digits_kernel<CurveBls, 0>")]
step = ls[starts[-1]:]
agg = collections.OrderedDict()
for name, ms, grid in step:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, (c, ms) in agg.items():
    print("%8.3f ms %4d  %s" % (ms, c, k))
print("%8.3f ms total over %d launches" % (tot, len(step)))
if len(sys.argv) > 2:
    for name, ms, grid in step:
        print("   %8.3f %s %s" % (ms, grid, name))
