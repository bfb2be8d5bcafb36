"""Count the Blackwell-specific SASS instructions of libwsb.so per kernel family (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UCGABAR_* = barrier.cluster, MAPA = mapa (DSMEM), HMMA = mma.sync."""
import collections
import os
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "whisperseg_b200", "libwsb.so")
WANT = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "HMMA")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.defaultdict(collections.Counter)
variants = collections.Counter()
fam = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        names = re.findall(r"([A-Za-z_][A-Za-z0-9_]*_kernel)", dem)
        fam = names[0] if names else dem[:40]
        variants[fam] += 1
        continue
    if fam is None:
        continue
    for tok in re.findall(r"\b([A-Z][A-Z0-9_]+)(?:\.[A-Z0-9_.]+)?\b", line):
        if tok in WANT:
            counts[fam][tok] += 1
print("%-34s %4s  %s" % ("kernel family", "inst", "instruction counts summed over the template instances"))
for fam in sorted(counts, key=lambda f: -sum(counts[f].values())):
    print("%-34s %4d  %s" % (fam, variants[fam], "  ".join("%s=%d" % kv for kv in sorted(counts[fam].items()))))
