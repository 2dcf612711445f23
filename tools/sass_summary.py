"""Per-kernel counts of the SASS mnemonics that prove what hardware path a kernel uses
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP / UTMALDG = bulk / tensor TMA copies,
LDGSTS = cp.async, UCGABAR = cluster barrier, FFMA2 / FMUL2 = packed FP32, DFMA = FP64) from the built library:
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepvcp-pointcloud-registration_b200", "libdvcp_b200.so")
WATCH = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "UCGABAR", "FFMA2", "FMUL2", "FFMA", "DFMA",
         "SHFL", "REDUX", "VOTE", "ATOMS", "LDS", "STS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)   # template-argument casts
            cur = per.setdefault(re.sub(r"\(.*", "", name).replace("void ", ""), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    print("SASS summary of %s (cuobjdump -sass; instruction counts per kernel, static)" % os.path.relpath(LIB, ROOT))
    print("%-58s %7s " % ("kernel", "total") + " ".join("%7s" % w for w in WATCH))
    for name, c in per.items():
        print("%-58s %7d " % (name[:58], c["total"]) + " ".join("%7d" % c[w] for w in WATCH))


if __name__ == "__main__":
    main()
