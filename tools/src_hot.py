"""Hot source lines per kernel from
   ncu -i X.ncu-rep --page source --print-source sass,cuda --csv > src.csv
usage: python tools/src_hot.py src.csv <function-substring> [top] [occurrence]"""
import collections
import csv
import sys

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = csv.reader(open(path))
func, fpath, hdr, seen = None, None, None, -1
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
        continue
    if r[0] == "Function Name":
        if func != r[1] or True:
            pass
        if r[1] != func:
            func = r[1]
            if want in func:
                seen += 1
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if func is None or want not in func or seen != occ or hdr is None:
        continue
    try:
        ie = int(r[hdr["Instructions Executed"]])
        te = int(r[hdr["Thread Instructions Executed"]])
        sm = int(r[hdr["# Samples"]])
    except (ValueError, IndexError):
        continue
    key = (fpath.split("/")[-1], r[0])
    a = agg[key]
    a[0] += ie
    a[1] += te
    a[2] += sm
    a[3] = r[1].strip()[:100]
tot = sum(a[0] for a in agg.values())
smp = sum(a[2] for a in agg.values())
print("function ~", want, "| warp instructions", tot, "| samples", smp)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %5s %6.2f%% inst %5.1f thr/inst %5.1f%% smpl | %s" % (key[0][:18], key[1], 100.0 * a[0] / max(tot, 1), a[1] / max(a[0], 1), 100.0 * a[2] / max(smp, 1), a[3]))
