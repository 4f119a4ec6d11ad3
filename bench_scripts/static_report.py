#!/usr/bin/env python
"""static_report.py — what the compiler made of libt4k.so, without a GPU (B200_PROFILING.md: "check -Xptxas -v and cuobjdump -sass
before spending GPU time").  Per kernel of the step / GEMM / conv / exchange paths: registers, spill, static shared memory
(`cuobjdump -res-usage`) and the count of the Blackwell-only SASS opcodes that prove which hardware path the kernel takes:

  UTCHMMA / UTCQMMA  tcgen05.mma (kind::tf32 / f16)        UTCBAR     tcgen05.commit
  LDTM / STTM        tcgen05.ld / tcgen05.st (TMEM)        UTMALDG    cp.async.bulk.tensor (TMA tensor-map load)
  UBLKCP             cp.async.bulk (1-D bulk copy)         LDGSTS     cp.async (asynchronous global -> shared copy)
  SYNCS              mbarrier operations                   UCGABAR    cluster barrier          MEMBAR.SYS  fence.sys (peer exchange)

usage: python bench_scripts/static_report.py [libt4k.so] > profiles/rNN_static_sass_report.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tensorforth_b200", "libt4k.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "UCGABAR", "MEMBAR.SC.SYS", "HMMA", "FFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(n):
    n = re.sub(r"\(.*$", "", n)                      # drop the parameter list
    n = re.sub(r"^void\s+", "", n)
    n = n.replace("t4k::", "").replace("(anonymous namespace)::", "")
    return n if len(n) <= 70 else n[:67] + "..."


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    hist, cur = collections.defaultdict(collections.Counter), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            hist[cur]["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    hist[cur][o] += 1
    names = demangle(sorted(usage))
    rows = []
    for k, (reg, stack, shared, local) in usage.items():
        h = hist.get(k, {})
        rows.append((short(names[k]), reg, stack, shared, h.get("_total", 0), [h.get(o, 0) for o in OPS]))
    rows.sort(key=lambda r: r[0])
    print("libt4k.so: %d kernels, all sm_100a (cuobjdump -res-usage / -sass; no GPU involved)" % len(rows))
    print("kernels with a stack frame (STACK > 0: local arrays or spills): %s" % (", ".join("%s (%d B)" % (r[0], r[2]) for r in rows if r[2]) or "none"))
    tot = collections.Counter()
    for r in rows:
        for o, v in zip(OPS, r[5]):
            tot[o] += v
    print("opcode totals: " + "  ".join("%s %d" % (o, tot[o]) for o in OPS))
    print()
    hdr = "%-70s %4s %5s %6s %6s  " % ("kernel", "regs", "stack", "smem", "instr") + " ".join("%7s" % o[:7] for o in OPS)
    print(hdr)
    print("-" * len(hdr))
    for name, reg, stack, shared, total, ops in rows:
        print("%-70s %4d %5d %6d %6d  " % (name, reg, stack, shared, total) + " ".join("%7s" % (v or ".") for v in ops))


if __name__ == "__main__":
    main()
