"""Opcode histogram per kernel of libgpmdm_sm100a.so (cuobjdump -sass): the SASS evidence for the Blackwell-native paths --
DMMA (fp64 tensor MMA), UBLKCP (TMA bulk copy), SYNCS (mbarrier), UTCHMMA / UTCBAR / LDTM (tcgen05 MMA, commit, TMEM load).
    python tools/sass_histogram.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gpmdm_b200", "lib", "libgpmdm_sm100a.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "LDS", "STS",
       "LDG", "STG", "ATOMG", "RED", "SHFL", "BAR", "FFMA2", "FADD2", "HMMA"]
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# cuobjdump -sass gpmdm_b200/lib/libgpmdm_sm100a.so (sm_100a), opcode counts per kernel (static instruction counts)")
print("# columns: total | " + " ".join(KEY))
tot = collections.Counter()
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    tot.update(h)
    name = demangle(k)
    cut = name.find(">(")
    name = (name[:cut + 1] if cut >= 0 else name.split("(")[0]).replace("void ", "").replace("gpmdm::", "").replace("(int)", "").replace("(bool)", "")
    print(f"{sum(h.values()):7d} | " + " ".join(f"{kk}={h[kk]}" for kk in KEY if h[kk]) + f"  :: {name}")
print("# library totals: " + " ".join(f"{kk}={tot[kk]}" for kk in KEY if tot[kk]))
