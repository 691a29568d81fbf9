"""Per-kernel time shares and DRAM bytes of the last complete subframe in an ncu launch list CSV
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum[,smsp__inst_executed.sum] --csv).
usage: python tools/launch_shares.py launches.csv out.json"""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
byid = {}
for r in rows[1:]:
    d = byid.setdefault(int(r[ii]), {"k": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(byid)
gen = [i for i in ids if "k_generate" in byid[i]["k"]]
res = [i for i in ids if "k_resolve" in byid[i]["k"]]
# the last FULL-SIZE subframe: bench.py ends with smaller diagnostic subframes (the counted one of the -DRT3_STATS twin, the
# CPU same-seed check); k_generate writes a fixed number of bytes per path, so its duration tells the size
full = max(byid[i]["gpu__time_duration.sum"] for i in gen)
last_gen = max(i for i in gen if byid[i]["gpu__time_duration.sum"] >= 0.8 * full and any(r > i for r in res))
last_res = min(r for r in res if r > last_gen)
sub = [byid[i] for i in ids if last_gen <= i <= last_res]
tot = sum(d["gpu__time_duration.sum"] for d in sub)
agg = {}
for d in sub:
    name = d["k"].split("(")[0].replace("void ", "").replace("rt3::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
    a[2] += d.get("dram__bytes_read.sum", 0)
    a[3] += d.get("dram__bytes_write.sum", 0)
    a[4] += d.get("smsp__inst_executed.sum", 0)
out = {"source": sys.argv[1], "total_ms_under_ncu": tot / 1e6, "kernels": {}}
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-28s launches %2d  %8.3f ms  share %5.1f%%  dram read %7.1f MB write %7.1f MB" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
    out["kernels"][k] = {"launches": a[0], "ms": a[1] / 1e6, "share": a[1] / tot, "dram_read_MB": a[2] / 1e6, "dram_write_MB": a[3] / 1e6}
    if a[4] > 0:
        out["kernels"][k]["warp_inst"] = a[4]   # warp-level instructions of these launches (smsp__inst_executed.sum)
json.dump(out, open(sys.argv[2], "w"), indent=1)
