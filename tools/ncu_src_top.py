#!/usr/bin/env python
"""Top source lines of an ncu report by warp-stall samples.

    ncu_src_top.py report.ncu-rep cubin kernel_substring [N]

ncu's CSV source page is per SASS instruction; nvdisasm -g gives the file:line of every instruction of the cubin
(built with -lineinfo).  The two listings are joined by instruction order inside each function (.text section).
Prints file:line (innermost inlined location), share of samples, the dominant stall reasons."""
import collections, csv, re, subprocess, sys
rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# sections in file order: name -> list of (offset, file:line, text)
secs, cur, loc = collections.OrderedDict(), None, "?"
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = m.group(1); secs[cur] = []; loc = "?"; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = m.group(1).split("/")[-1] + ":" + m.group(2); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m and cur:
        secs[cur].append((int(m.group(1), 16), loc, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ins = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
base = int(ins[0]["Address"], 16)
# ncu lists the kernel first, then the device functions it calls, each at its own address: match by opcode runs
flat = []
order = [k for k in secs if kname in k] + [k for k in secs if kname not in k]
for k in order:
    flat += [(k, o, l, t) for (o, l, t) in secs[k]]
def opc(t):
    t = t.strip()
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0] if t else ""
if len(flat) != len(ins):
    print("warning: %d disassembled vs %d profiled instructions; joining by order" % (len(flat), len(ins)))
bad = sum(1 for a, b in zip(flat, ins) if opc(a[3]) != opc(b["Source"]))
if bad:
    print("warning: %d opcode mismatches in the join" % bad)
def num(x):
    try: return float(x)
    except Exception: return 0.0
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.defaultdict(lambda: collections.Counter())
for a, b in zip(flat, ins):
    c = agg[a[2]]
    c["samples"] += num(b["# Samples"])
    c["inst"] += num(b["Instructions Executed"])
    for k in stalls:
        c[k] += num(b[k])
tot = sum(c["samples"] for c in agg.values())
print("total samples %.0f" % tot)
for loc, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    why = sorted(((c[k], k[6:]) for k in stalls), reverse=True)[:3]
    print("%5.1f%%  %-26s inst %10.0f | %s" % (100 * c["samples"] / tot, loc, c["inst"],
                                             ", ".join("%s %.0f%%" % (k, 100 * v / max(c["samples"], 1)) for v, k in why if v > 0)))
