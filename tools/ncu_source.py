"""Per-CUDA-source-line hot spots of one kernel launch in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_source.py file.ncu-rep <launch index> [topN]"""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# sections start with "Function Name"; inside, "File Path" rows switch files, data rows: line no, source, address, sass, metrics...
launches, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "files": {}, "hdr": None, "file": None}
        launches.append(cur)
    elif cur is None:
        continue
    elif r and r[0] == "File Path":
        cur["file"] = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["files"].setdefault(cur["file"], []).append(r)
L = launches[which]
h = L["hdr"]
ie, te, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
agg = {}
for f, rs in L["files"].items():
    line, text = None, ""
    for r in rs:
        if r[0]:
            line, text = r[0], r[1]
        k = (f, int(line) if line else -1)
        a = agg.setdefault(k, [0, 0, 0, text])
        try:
            a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm])
        except ValueError:
            pass
ti, ts = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
print(L["name"], "warp-inst", ti, "samples", ts)
for k, a in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print("%5.1f%% samp %5.1f%% inst thr/inst %4.1f  %s:%d | %s" % (100 * a[2] / max(ts, 1), 100 * a[0] / max(ti, 1), a[1] / max(a[0], 1), k[0], k[1], a[3].strip()[:100]))
