#!/usr/bin/env python3
"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch, --csv) -> per-kernel summary of the
LAST MSM step / NTT transform in the capture and profiles/r02_traffic.json (what bench.py reports as roofline.traffic).
usage: traffic_summary.py msm.csv ntt.csv out.json"""
import collections, csv, json, re, sys


def launches(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(rows):
        k = r["ID"]
        d = per.setdefault(k, {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("ab200::", "")})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ms"] = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            d[r["Metric Name"].split("__")[1].split(".")[0]] = v * mult
    return list(per.values())


def summarize(ls, first_kernel_prefix):
    starts = [i for i, l in enumerate(ls) if l["name"].startswith(first_kernel_prefix)]
    step = ls[starts[-1]:]
    agg = collections.OrderedDict()
    for l in step:
        a = agg.setdefault(l["name"], {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
        a["launches"] += 1
        a["ms"] += l.get("ms", 0.0)
        a["dram_bytes"] += l.get("bytes_read", 0.0) + l.get("bytes_write", 0.0)
    return step, agg


msm, ntt, out = sys.argv[1:4]
step, agg = summarize(launches(msm), "msm_digits_kernel<CurveBls, 0>")
dom = max(step, key=lambda l: l.get("ms", 0.0))
res = {
    "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on bench.py (n=2^26) / tools/ntt_time.py (n=2^24)",
    "msm_step_dram_bytes": sum(a["dram_bytes"] for a in agg.values()),
    "msm_step_ms_under_ncu": sum(a["ms"] for a in agg.values()),
    "msm_dominant_kernel": dom["name"],
    "msm_dominant_kernel_ms_under_ncu": dom.get("ms"),
    "msm_dominant_kernel_dram_bytes_per_launch": dom.get("bytes_read", 0.0) + dom.get("bytes_write", 0.0),
    "msm_kernels": agg,
}
nl = launches(ntt)
starts = [i for i, l in enumerate(nl) if "ntt" in l["name"] and "pass" in l["name"]]
# one transform = the last m consecutive pass launches
last = []
for l in reversed(nl):
    if "pass_kernel" in l["name"]:
        last.append(l)
        if len(last) == 3:
            break
res["ntt_passes"] = [{"name": l["name"], "ms": l.get("ms"), "dram_bytes": l.get("bytes_read", 0.0) + l.get("bytes_write", 0.0)} for l in reversed(last)]
res["ntt_dram_bytes_per_transform"] = sum(p["dram_bytes"] for p in res["ntt_passes"])
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k not in ("msm_kernels",)}, indent=1)[:1500])
for k, a in agg.items():
    print("%9.3f ms %8.2f GB x%-3d %s" % (a["ms"], a["dram_bytes"] / 1e9, a["launches"], k[:80]))
