#!/usr/bin/env python
"""Warp instructions and stall samples of one kernel of an ncu report, summed per source FUNCTION of one .cu file.

    ncu_phase_inst.py report.ncu-rep cubin kernel_substring source.cu

Every SASS instruction is attributed to the innermost frame of its inline chain (nvdisasm -gi) that lies in source.cu,
and that line to the function whose body contains it (so a DMMA issued from ldlt_device.cuh inside assemble_normal counts
for assemble_normal).  Joined with ncu's per-instruction source page by instruction order, as ncu_src_top.py does."""
import collections, csv, re, subprocess, sys
rep, cubin, kname, src = sys.argv[1:5]
base = src.split("/")[-1]
# function bodies of the source file: a line that starts a definition at column 0 up to the closing brace at column 0
funcs, cur = [], None
for i, ln in enumerate(open(src), 1):
    m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__)[^;(]*?\b(\w+)\s*\((?!FT)", ln)
    if m is None:
        m = re.match(r"^__global__.*\)\s+(\w+)\s*\(", ln)
    if m and cur is None and not ln.rstrip().endswith(";"):
        cur = [m.group(1), i, None]
    if ln.startswith("}") and cur is not None:
        cur[2] = i; funcs.append(tuple(cur)); cur = None
def func_of(line):
    for n, a, b in funcs:
        if a <= line <= b:
            return n
    return "?"
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
secs, cur, chain, fresh = collections.OrderedDict(), None, [], True
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = m.group(1); secs[cur] = []; chain = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # "-gi" prints one line per frame, innermost first; the frames of an instruction precede it
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and cur:
        inner = next((l for f, l in chain if f == base), None)
        secs[cur].append((func_of(inner) if inner else "?", m.group(2)))
        fresh = True
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ins = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
flat = []
for k in [k for k in secs if kname in k] + [k for k in secs if kname not in k]:
    flat += secs[k]
if len(flat) != len(ins):
    print("warning: %d disassembled vs %d profiled instructions" % (len(flat), len(ins)))
def num(x):
    try: return float(x)
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (fn, text), b in zip(flat, ins):
    a = agg[fn]
    a[0] += num(b["Instructions Executed"]); a[1] += num(b["# Samples"])
    if re.search(r"\bDMMA\b", text): a[2] += num(b["Instructions Executed"])
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("%-28s %14s %7s %9s %12s" % ("function", "warp instr", "share", "samples", "DMMA instr"))
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %14.0f %6.1f%% %8.1f%% %12.0f" % (fn, a[0], 100 * a[0] / ti, 100 * a[1] / ts, a[2]))
print("%-28s %14.0f" % ("total", ti))
