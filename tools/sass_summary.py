#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library: counts of the Blackwell-native instructions
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,
UTCBAR = tcgen05.commit, LDGSTS = cp.async, SYNCS = mbarrier) and of the legacy tensor path (HMMA) in every kernel of
gencast_flax_nnx_b200/libgencast_b200.so.  Runs on the build box (cuobjdump, no GPU):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gencast_flax_nnx_b200", "libgencast_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDGSTS", "SYNCS", "HMMA", "FFMA", "MUFU"]


def demangle(name: str) -> str:
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        out = name
    out = re.sub(r"gc::\(anonymous namespace\)::", "", out)
    return re.sub(r"\(.*\)$", "", out)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        base = op.split(".")[0]
        cur[base] += 1
        if op.startswith("UTCHMMA.2CTA"):
            cur["UTCHMMA.2CTA"] += 1
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; static instruction counts per kernel)")
    print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>9s}" for k in KEYS))
    for name, c in kernels.items():
        print(f"{demangle(name)[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[k]:9d}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
