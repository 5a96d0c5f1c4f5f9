"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library (evidence that the hot kernels are
tcgen05 / TMEM / TMA code): python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flocoder_b200", "_C", "libflocoder_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "MUFU", "SHFL", "BAR.SYNC", "BAR.ARV"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, regs, cur = collections.OrderedDict(), {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    counts[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line) else 0
    for mn in MNEMONICS:
        if re.search(r"\b" + re.escape(mn), line):
            counts[cur][mn] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); counts of instruction mnemonics per kernel")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk,")
print("# LDTM / STTM = tcgen05.ld / st (tensor memory), SYNCS = mbarrier ops")
hdr = ["kernel"] + ["instructions"] + MNEMONICS
print("\t".join(hdr))
for k, c in counts.items():
    if not any(c[m] for m in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM")) and "--all" not in sys.argv:
        continue
    print("\t".join([k] + [str(c[h]) for h in hdr[1:]]))
