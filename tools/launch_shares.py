"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ns and share per kernel.
usage: python tools/launch_shares.py gpurun_out/launches.csv profiles/r1_launch_shares.json"""
import csv, json, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = {}
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    a = agg.setdefault(name, {"launches": 0, "ns_total": 0.0})
    a["launches"] += 1
    a["ns_total"] += float(r[vi].replace(",", ""))
tot = sum(a["ns_total"] for a in agg.values())
for a in agg.values():
    a["share"] = a["ns_total"] / tot
out = dict(sorted(agg.items(), key=lambda kv: -kv[1]["ns_total"]))
json.dump({"total_ns": tot, "kernels": out}, open(sys.argv[2], "w"), indent=1)
for k, a in list(out.items())[:8]:
    print("%6.2f %%  %5d  %s" % (100 * a["share"], a["launches"], k[:90]))
