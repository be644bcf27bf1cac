"""Mnemonic counts per kernel from `cuobjdump -sass vstrains_b200/libvspe.so` (evidence that the bulk-TMA / mbarrier /
warp-vote paths are in the built library).  usage: python tools/sass_summary.py > profiles/r02_sass_mnemonics.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vstrains_b200", "libvspe.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print("# cuobjdump -sass vstrains_b200/libvspe.so @ %s -- mnemonic counts per kernel" % head)
print(elf.strip())
pats = {"UBLKCP": r"\bUBLKCP", "SYNCS.ARRIVE.TRANS": r"SYNCS\.ARRIVE\.TRANS", "SYNCS.PHASECHK": r"SYNCS\.PHASECHK", "VOTE": r"\bVOTE\.",
        "MATCH": r"\bMATCH\.", "REDUX": r"\bREDUX", "SHFL": r"\bSHFL\.", "ATOM": r"\bATOM", "RED": r"\bRED\.", "LDG.E.128": r"LDG\.E\.128",
        "STG.E.128": r"STG\.E\.128", "LDS.128": r"LDS\.128", "BAR.SYNC": r"BAR\.SYNC", "UTMALDG": r"UTMALDG", "UTCMMA": r"UTCMMA"}
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        continue
    cnt[cur]["insts"] += 1
    for k, p in pats.items():
        if re.search(p, line):
            cnt[cur][k] += 1
for f, c in sorted(cnt.items(), key=lambda kv: -kv[1]["insts"]):
    name = re.sub(r"\(.*", "", subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip())
    print("%-52s insts=%-6d %s" % (name[:52], c["insts"], " ".join("%s=%d" % (k, v) for k, v in c.items() if k != "insts")))
