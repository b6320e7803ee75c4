#!/usr/bin/env python
"""Static SASS mnemonic counts per kernel of the built library -> profiles/sass_summary.txt

    python scripts/sass_summary.py > profiles/sass_summary.txt

Evidence that the hot kernels really are tcgen05 / TMA / TMEM code (UTCHMMA, UTMALDG, LDTM, UTCBAR; the `.2CTA`
forms are the cta_group::2 instructions of the CTA-pair mode of k_gemm_gnt) and which ones use mma.sync (HMMA).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffuscene_b200", "lib", "libdiffuscene_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG", "UTCBAR.2CTA", "UTCBAR", "LDTM", "HMMA", "LDSM", "STSM",
        "SYNCS", "REDG", "MUFU.TANH", "MUFU.EX2", "ELECT", "R2UR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    cur, i = None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[i].replace("(anonymous namespace)::", ""))
            i += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                break
    print("SASS mnemonic counts per kernel of diffuscene_b200/lib/libdiffuscene_b200.so (cuobjdump -sass, sm_100a; static counts).")
    print("UTCHMMA = tcgen05.mma (bf16), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, .2CTA = the")
    print("cta_group::2 forms (CTA-pair mode), HMMA = mma.sync, LDSM / STSM = ldmatrix / stmatrix, SYNCS = mbarrier ops,")
    print("REDG = red.global (vector fp32 atomics of the split-K dW GEMMs), ELECT / R2UR = lane election / vector->uniform moves")
    print("of the single-lane roles.\n")
    rows = [(n, c) for n, c in counts.items() if any(c[k] for k in ("UTCHMMA", "UTCHMMA.2CTA", "HMMA", "LDTM"))]
    rows.sort(key=lambda r: -(r[1]["UTCHMMA"] + r[1]["UTCHMMA.2CTA"]) * 1000 - r[1]["HMMA"])
    for n, c in rows:
        print("%-46s %s" % (n[:46], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))


if __name__ == "__main__":
    sys.exit(main())
