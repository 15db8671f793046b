#!/usr/bin/env python
"""Per-kernel SASS / resource evidence from the built library (no GPU needed): counts of the tcgen05 / TMA /
TMEM mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, TMA -> UTMALDG/UTMASTG, tcgen05.ld/st -> LDTM/STTM),
registers, spills and static shared memory per kernel.  Usage: python tools/sass_evidence.py > profiles/<file>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ace-step-1.5-for-windows_b200", "libacestep_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS",
             "ELECT", "MUFU.EX2", "MUFU.RCP", "HMMA", "FFMA", "ACQBULK", "REDUX"]


def short(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    out = out.replace("ace::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    out = re.sub(r"\(.*$", "", out)
    return out[:78]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_instr"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    cur[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)), re.search(r"LOCAL:(\d+)", line))
    print("# SASS evidence per kernel of libacestep_b200.so (cuobjdump -sass / -res-usage, sm_100a); counts are static")
    print("# instructions.  UTCHMMA = tcgen05.mma (…2CTA = cta_group::2), UTMALDG = TMA tensor load, LDTM/STTM =")
    print("# tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, MUFU.EX2 = ex2.approx.")
    print(f"{'kernel':78s} {'instr':>6s} {'regs':>4s} {'smem':>6s} {'local':>5s}  mnemonics")
    for name, c in kernels.items():
        reg, smem, loc = usage.get(name, ("?", "?", None))
        loc = loc.group(1) if loc else "?"
        mn = " ".join(f"{k}={c[k]}" for k in MNEMONICS + ["UTCHMMA.2CTA"] if c[k] and k not in ("FFMA",))
        print(f"{short(name):78s} {c['_instr']:6d} {reg!s:>4s} {smem!s:>6s} {loc:>5s}  {mn}")


if __name__ == "__main__":
    sys.exit(main())
