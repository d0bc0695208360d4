"""Per-kernel SASS opcode histogram of libiaf_b200.so: the instructions that prove which hardware paths the kernels use
(tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, tcgen05.cp = UTCCP, cp.async.bulk = UBLKCP, TMA tensor = UTMALDG,
mbarrier = SYNCS, tcgen05.commit = UTCBAR, FFMA for the SIMT kernels).  Usage: python tools/sass_histogram.py > profiles/rN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "iaf_b200", "lib", "libiaf_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCCP", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "HMMA", "FFMA",
        "FFMA2", "FADD2", "FMUL2", "MUFU", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
            kernels[cur]["__all__"] += 1
    dem = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, dem))
    # one representative instantiation per kernel template: the one the headline workloads launch when present
    groups = collections.OrderedDict()
    for k in kernels:
        base = re.sub(r"<.*", "", names[k]).replace("void ", "")
        groups.setdefault(base, []).append(k)
    print("# SASS opcode histogram of `iaf_b200/lib/libiaf_b200.so` (sm_100a)\n")
    print("`cuobjdump -sass`, static instruction counts per kernel.  `UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st,")
    print("`UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (TMA bulk copy), `SYNCS` = mbarrier ops.  Template kernels: the")
    print("row is the instantiation with the most instructions of that family; `n` = number of instantiations in the library.\n")
    cols = [c for c in KEYS if any(kernels[k][c] for k in kernels)]
    print("| kernel | n | instrs | " + " | ".join(cols) + " |")
    print("|---|---:|---:|" + "---:|" * len(cols))
    for base, ks in groups.items():
        k = max(ks, key=lambda x: kernels[x]["__all__"])
        print("| `%s` | %d | %d | " % (base, len(ks), kernels[k]["__all__"]) + " | ".join(str(kernels[k][c] or "") for c in cols) + " |")
    tot = collections.Counter()
    for k in kernels:
        tot.update(kernels[k])
    print("\nWhole library: " + ", ".join("%s %d" % (c, tot[c]) for c in cols if tot[c]))


if __name__ == "__main__":
    sys.exit(main())
