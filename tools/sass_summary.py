"""SASS evidence for profiles/: per kernel of libmtts.so, the counts of the tcgen05 / TMA / TMEM instructions
(UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,
UTMASTG = TMA store) from `cuobjdump -sass`, plus registers per thread from `cuobjdump -res-usage`.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "meta-tts_b200", "libmtts.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "MUFU", "STG", "LDG", "RED", "ATOMG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    cur[mn] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and fn:
            regs[fn] = int(m.group(1))
    dm = demangle(list(kernels))
    short = lambda n: re.sub(r"\(anonymous namespace\)::|void |<unnamed>::", "", dm.get(n, n)).split("(")[0]  # noqa: E731
    print("# SASS summary of meta-tts_b200/libmtts.so (sm_100a; `python tools/sass_summary.py`)\n")
    print("Counts of static SASS instructions per kernel.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load (cp.async.bulk.tensor), "
          "LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations.  A kernel with UTCHMMA > 0 "
          "computes on the 5th-generation tensor cores with accumulators in TMEM; none of them uses mma.sync / wgmma.\n")
    tc = [(n, c) for n, c in kernels.items() if c["UTCHMMA"]]
    print(f"Tensor-core kernels: {len(tc)} of {len(kernels)}; library totals: " +
          ", ".join(f"{mn} {sum(c[mn] for c in kernels.values())}" for mn in MNEMONICS[:6]) + "\n")
    print("| kernel | regs | instr | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|---|" + "---|" * len(MNEMONICS))
    for n, c in sorted(kernels.items(), key=lambda kv: (-kv[1]["UTCHMMA"], short(kv[0]))):
        print(f"| `{short(n)}` | {regs.get(n, '')} | {c['_total']} | " + " | ".join(str(c[mn]) if c[mn] else "" for mn in MNEMONICS) + " |")
    hm = sum(1 for line in sass.splitlines() if re.search(r"\sHMMA|\sHGMMA|\sWGMMA", line))
    print(f"\nLegacy tensor instructions (HMMA / HGMMA): {hm}.")


if __name__ == "__main__":
    sys.exit(main())
