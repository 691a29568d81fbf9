"""Per-CUDA-source-line hot spots of one kernel launch in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_source.py file.ncu-rep <launch index> [topN]"""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the page is a sequence of (File Path, Function Name, header, rows...) blocks, one per source file per launch
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
    line, text = None, ""
    for r in b["rows"]:
        if r[0]:
            line, text = int(r[0]), r[1]
        a = agg.setdefault((b["file"], line), [0, 0, 0, text, 0])
        try:
            a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm]); a[4] += 1
        except ValueError:
            pass
ti, ts = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
print(launches[which][0]["name"], "launches in report:", len(launches), "| warp-inst", ti, "samples", ts)
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp sass %3d thr/inst %4.1f  %s:%s | %s" % (100 * a[0] / max(ti, 1), 100 * a[2] / max(ts, 1), a[4], a[1] / max(a[0], 1), k[0][4:12], k[1], a[3].strip()[:96]))
