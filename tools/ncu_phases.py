"""Warp-instruction share per source FUNCTION of one kernel launch in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_phases.py file.ncu-rep [launch index]"""
import csv
import os
import re
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rendertoy3c_b200", "csrc")
fn_of = {}
for f in os.listdir(root):
    cur, table = "?", []
    for i, line in enumerate(open(os.path.join(root, f)), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:RT3_HD|__device__ __forceinline__|__global__|static|RT3_GLOBAL)\b[^;(]*?\b(\w+)\s*\(", line)
        if m and not line.strip().startswith("//"):
            cur = m.group(1)
        table.append(cur)
    fn_of[f] = table
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur, lastfile = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        lastfile = r[1].split("/")[-1]
    elif r and r[0] == "Function Name":
        cur = {"name": r[1], "file": lastfile, "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is None:
        continue
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
launches, seen = [], set()
for b in blocks:
    if not launches or b["file"] in seen:
        launches.append([])
        seen = set()
    launches[-1].append(b)
    seen.add(b["file"])
agg = {}
for b in launches[which]:
    h = b["hdr"]
    ie, te, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    line = None
    for r in b["rows"]:
        if r[0]:
            line = int(r[0])
        t = fn_of.get(b["file"])
        g = (t[line - 1] if t and line and line <= len(t) else b["file"])
        a = agg.setdefault(g, [0, 0, 0])
        try:
            a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm])
        except ValueError:
            pass
ti, ts = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
print(launches[which][0]["name"], "| launches in report:", len(launches), "| warp-inst", ti)
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:24]:
    print("%-28s %5.1f%% inst %5.1f%% samples  %4.1f threads/inst" % (k, 100 * a[0] / ti, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
