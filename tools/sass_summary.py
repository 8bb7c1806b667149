#!/usr/bin/env python3
"""Per-kernel SASS evidence of libdnaldpc.so (sm_100a cubins): instruction counts that show what each kernel is built
from - bulk copies (UBLKCP = cp.async.bulk, the TMA engine), mbarrier ops (SYNCS), tensor-memory stores / loads
(STTM / LDTM = tcgen05.st / ld), tensor-memory allocation (UTCATOMSWS), async copies (LDGSTS), the fp64 reciprocal seed
(MUFU.RCP64H) and the fp64 pipe instructions - plus registers per thread from cuobjdump --dump-resource-usage.
usage: sass_summary.py [path/to/libdnaldpc.so] > profiles/rNN/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UBLKCP", "SYNCS", "STTM", "LDTM", "UTCATOMSWS", "LDGSTS", "MUFU.RCP64H", "MUFU", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "ATOMG", "SHFL", "VOTE", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dna-ldpc-codes_b200", "libdnaldpc.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    names = demangle(list(counts))
    print("# SASS summary of %s\n" % os.path.relpath(lib, ROOT))
    print("Architectures in the fat binary: %s (cuobjdump -sass). Counts are static instructions per kernel; `MUFU` includes `MUFU.RCP64H`.\n" % ", ".join(sorted(archs)))
    print("| kernel | regs | static smem | instrs | " + " | ".join(OPS) + " |")
    print("|---|---|---|---|" + "---|" * len(OPS))
    for k, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
        n = names[k].replace("dnaldpc::", "").replace("void ", "")
        n = re.sub(r"\(.*", "", n)
        r = regs.get(k, ("?", "?"))
        print("| %s | %s | %s | %d | " % (n, r[0], r[1], c["_total"]) + " | ".join(str(c[o]) for o in OPS) + " |")


if __name__ == "__main__":
    main()
