#!/usr/bin/env python
"""SASS census of libspacer_b200's objects: per translation unit and per kernel, how many tcgen05 MMA (UTCHMMA...), TMA
(UTMALDG/UTMASTG/UBLKCP), TMEM load/store (LDTM/STTM) and legacy mma.sync (HMMA) instructions the sm_100a code holds.
    python tools/sass_census.py > profiles/r02_sass_census.md        (needs only cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "spacer_b200", "lib", "obj")
CLASSES = [("tcgen05.mma", r"\bUTC[A-Z]*MMA"), ("tcgen05.cp/shift", r"\bUTCCP|\bUTCSHIFT"), ("TMA load", r"\bUTMALDG"),
           ("TMA store", r"\bUTMASTG|\bUTMAREDG"), ("bulk copy / L2 prefetch", r"\bUBLKCP|\bUBLKPF|\bUTMAPF"),
           ("TMEM ld", r"\bLDTM"), ("TMEM st", r"\bSTTM"), ("TMEM alloc", r"\bUTCATOM|\bUTCBAR|\bUTCALLOC"),
           ("mbarrier", r"\bSYNCS"), ("mma.sync HMMA", r"\bHMMA"), ("cluster DSMEM", r"\bUCGABAR|\bCCTL|\bMAPA")]


def census(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for name, rx in CLASSES:
            if re.search(rx, line):
                per[cur][name] += 1
    return per


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    print("# SASS census of the sm_100a objects (cuobjdump -sass, CUDA 12.9) -- round 2\n")
    print("Counts are static instruction counts per kernel.  `UTC*MMA` = tcgen05.mma, `UTMALDG` = TMA tensor load, `LDTM`/`STTM`"
          " = tcgen05.ld/st (TMEM), `HMMA` = legacy mma.sync.\n")
    heads = [c[0] for c in CLASSES]
    total_by_obj = []
    detail = []
    for fn in sorted(os.listdir(OBJ)):
        if not fn.endswith(".o"):
            continue
        per = census(os.path.join(OBJ, fn))
        tot = collections.Counter()
        for c in per.values():
            tot.update(c)
        total_by_obj.append((fn, tot, len(per)))
        dm = demangle(list(per))
        for k, c in per.items():
            if any(c[h] for h in ("tcgen05.mma", "TMA load", "TMEM ld", "TMEM st", "mma.sync HMMA")):
                short = dm[k].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
                short = re.sub(r"^void ", "", short)
                short = re.sub(r">\(.*$", ">", short)            # drop the parameter list, keep the template arguments
                detail.append((fn, short[:120], c))
    print("## Per object\n")
    print("| object | kernels | " + " | ".join(heads) + " |")
    print("|---|---|" + "---|" * len(heads))
    for fn, tot, n in total_by_obj:
        print(f"| {fn} | {n} | " + " | ".join(str(tot[h]) for h in heads) + " |")
    print("\n## Kernels that use tensor cores, TMA or TMEM\n")
    cols = ["tcgen05.mma", "TMA load", "TMEM ld", "TMEM st", "mma.sync HMMA"]
    print("| object | kernel | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for fn, k, c in detail:
        print(f"| {fn} | `{k}` | " + " | ".join(str(c[h]) for h in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
