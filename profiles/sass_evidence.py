#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel family the number of template instantiations, the
range of registers / static shared memory / spills (`cuobjdump --dump-resource-usage`) and which SASS mnemonics of
interest the family contains (`cuobjdump -sass`): UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, REDUX = warp reduction
in the uniform datapath, LDG/STG.128 = 128-bit global accesses, BRX = jump table, CCTL / prefetch = L2 prefetch.

    python profiles/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "awkward-0.x_b200", "lib", "libjagged_sm100a.so")

MNEMONICS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UTMALDG", r"\bUTMALDG"), ("REDUX", r"\bREDUX"),
             ("LDG.128", r"\bLDG\.E[.\w]*\.128"), ("STG.128", r"\bSTG\.E[.\w]*\.128"), ("LDS.128", r"\bLDS[.\w]*\.128"),
             ("BRX", r"\bBRX"), ("CCTL", r"\bCCTL"), ("ATOM/RED", r"\b(ATOMG|RED)\."), ("SHFL", r"\bSHFL"),
             ("VOTE", r"\bVOTE"), ("NANOSLEEP", r"\bNANOSLEEP")]


def family(mangled):
    out = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"^void\s+", "", out)
    m = re.match(r"([\w:]+)", out)
    return m.group(1) if m else mangled


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first: python -c 'import __graft_entry__ as g; g.build()'")
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            f = dict(kv.split(":") for kv in line.split() if ":" in kv)
            usage[name] = (int(f.get("REG", 0)), int(f.get("SHARED", 0)), int(f.get("STACK", 0)), int(f.get("LOCAL", 0)))
            name = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    cur = None
    pats = [(k, re.compile(p)) for k, p in MNEMONICS]
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4}\*/", line):
            ninstr[cur] += 1
            for k, p in pats:
                if p.search(line):
                    per[cur][k] += 1
    names = sorted(set(usage) | set(ninstr))
    cache = {}
    fams = collections.defaultdict(list)
    for n in names:
        fams[cache.setdefault(n, family(n))].append(n)
    total = collections.Counter()
    for n in names:
        total.update(per[n])
    print("# static evidence from awkward-0.x_b200/lib/libjagged_sm100a.so (sm_100a cubins; profiles/sass_evidence.py)")
    print("# device functions: %d in %d kernel families; SASS instructions: %d" % (len(names), len(fams), sum(ninstr.values())))
    print("# mnemonic sites in the whole library: " + ", ".join("%s %d" % (k, total[k]) for k, _ in MNEMONICS))
    print("# UBLKCP = cp.async.bulk.shared::cluster.global.mbarrier (1-D TMA bulk copy), SYNCS = mbarrier init / expect_tx /")
    print("# try_wait; no UTMALDG: every staged run is 1-D contiguous (no tensor map); no tensor-core mnemonics: nothing is a contraction")
    print()
    hdr = "%-46s %5s %9s %13s %7s %9s  %s" % ("kernel family", "inst.", "registers", "static smem B", "spill B", "SASS/inst", "mnemonics (sites per instantiation, max)")
    print(hdr)
    print("-" * len(hdr))
    for fam in sorted(fams, key=lambda f: -sum(ninstr[n] for n in fams[f])):
        ns = fams[fam]
        regs = [usage[n][0] for n in ns if n in usage]
        sh = [usage[n][1] for n in ns if n in usage]
        spill = [max(usage[n][2], usage[n][3]) for n in ns if n in usage]
        marks = []
        for k, _ in MNEMONICS:
            mx = max(per[n][k] for n in ns)
            if mx:
                marks.append("%s %d" % (k, mx))
        rng = lambda v: ("%d" % v[0] if min(v) == max(v) else "%d-%d" % (min(v), max(v))) if v else "-"
        print("%-46s %5d %9s %13s %7s %9d  %s" % (fam[:46], len(ns), rng(regs), rng(sh), rng(spill),
                                                 max(ninstr[n] for n in ns), ", ".join(marks)))


if __name__ == "__main__":
    main()
