#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics per kernel of the built libemd.so (cuobjdump -sass): the evidence that the hot
path is tcgen05 / TMEM / TMA code (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG /
UBLKCP; HMMA would be the legacy mma.sync path).   python tools/sass_summary.py [--out profiles/x.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ai-cv-automation-elect-micr_b200", "libemd.so")
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA2", "HMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    cur[w] += 1
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "HMMA", "LDGSTS", "_total"]
    lines = ["# SASS mnemonic counts per kernel of libemd.so (sm_100a), cuobjdump -sass; UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), "
             "LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, HMMA = legacy mma.sync (none expected)",
             f"{'kernel':90s} " + " ".join(f"{c:>12s}" for c in cols)]
    tot = collections.Counter()
    for name, c in per.items():
        if not any(c[k] for k in cols[:6]):
            continue
        lines.append(f"{name[:90]:90s} " + " ".join(f"{c[k]:12d}" for k in cols))
        tot.update(c)
    lines.append(f"{'TOTAL (kernels with tensor-core / TMA instructions)':90s} " + " ".join(f"{tot[k]:12d}" for k in cols))
    others = [n for n, c in per.items() if not any(c[k] for k in cols[:6])]
    lines.append(f"# {len(others)} other kernels (CUDA-core: FP32 validation mode, stem, resize, pool, wrapper, quality): " + ", ".join(sorted(set(o[:40] for o in others)))[:1500])
    text = "\n".join(lines)
    print(text)
    if "--out" in sys.argv:
        open(os.path.join(ROOT, sys.argv[sys.argv.index("--out") + 1]), "w").write(text + "\n")


if __name__ == "__main__":
    main()
