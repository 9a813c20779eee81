#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware paths the shipped library uses:
DMMA (FP64 tensor pipe), UTMALDG (TMA tensor loads), UBLKCP (TMA bulk copies), SYNCS (mbarrier), LDGSTS (cp.async),
BAR, DFMA, and the register count.  Usage: sass_extract.py ipm-zoo_b200/libipmz_b200.so > profiles/rNN_sass_extract.txt"""
import collections, re, subprocess, sys
so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
name = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m: name = m.group(1); continue
    m = re.search(r"REG:(\d+)", ln)
    if m and name: regs[name] = int(m.group(1)); name = None
want = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "BAR", "DFMA", "SHFL", "STL", "LDL"]
cnt, cur, arch = collections.OrderedDict(), None, "?"
for ln in sass.splitlines():
    m = re.search(r"arch = (sm_\w+)", ln)
    if m: arch = m.group(1)
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); cnt[cur]["arch"] = arch; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        cnt[cur]["instr"] += 1
        if op in want: cnt[cur][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("%s: %d kernels / device functions" % (so, len(cnt)))
print("%-58s %-8s %5s %7s " % ("kernel", "arch", "regs", "instr") + " ".join("%7s" % w for w in want))
for (k, c), d in zip(cnt.items(), dem):
    short = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("ipmz::", ""))
    print("%-58s %-8s %5s %7d " % (short[-58:], c["arch"], regs.get(k, ""), c["instr"]) + " ".join("%7d" % c[w] for w in want))
tot = collections.Counter()
for c in cnt.values():
    for w in want: tot[w] += c[w]
print("%-58s %-8s %5s %7s " % ("total", "", "", "") + " ".join("%7d" % tot[w] for w in want))
