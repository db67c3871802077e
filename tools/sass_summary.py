"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA, TMEM loads, TMA / bulk copies)
from `cuobjdump -sass` of the in-tree libdcap.so -> profiles/sass_summary.txt.  Runs anywhere (no GPU needed):
    python tools/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "image-captioning_b200", "libdcap.so")
OUT = os.path.join(ROOT, "profiles", "sass_summary.txt")
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "UTCCP", "UTMALDG.2D.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG",
       "UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK", "LDS.128", "LDG.E.128", "STG.E.EF", "RED.E", "REDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = counts.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PAT:                      # longest patterns first in PAT where one is a prefix of another
            if op.startswith(p):
                cur[p] += 1
                break
    with open(OUT, "w") as f:
        f.write("# cuobjdump -sass image-captioning_b200/libdcap.so: per-kernel counts of the Blackwell-specific mnemonics\n")
        f.write("# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce, UBLKCP = cp.async.bulk,\n")
        f.write("#  SYNCS.* = mbarrier transactions).  Kernels without any of them are omitted.  Regenerate: python tools/sass_summary.py\n")
        tot = collections.Counter()
        for name, c in counts.items():
            if not c:
                continue
            tot.update(c)
            f.write("%s\n    %s\n" % (name, "  ".join("%s x%d" % (k, v) for k, v in sorted(c.items()))))
        f.write("TOTAL\n    %s\n" % "  ".join("%s x%d" % (k, v) for k, v in sorted(tot.items())))
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
