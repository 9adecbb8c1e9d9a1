"""Top source lines of a kernel by executed instructions / stall samples, attributed to the outermost frame inside the
model files (tools, not product).  usage: ncu_lines.py <nvdisasm -gi dump> <mangled kernel> <ncu source csv> [n]"""
import collections
import csv
import re
import sys

disasm, func, srccsv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section") and (".text." + func) in l)
chain, off2, pending = [], {}, True
for l in lines[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        if pending:
            chain, pending = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?(\S+)", l)
    if m:
        off2[int(m.group(1), 16)] = (tuple(chain), m.group(2).rstrip(";"))
        pending = True
rows = list(csv.reader(open(srccsv)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), int(r[ie]), int(r[isamp])))
    except (ValueError, IndexError):
        pass
base = data[0][0]
byl, sm, byf, smf = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot, ts = sum(e for _, e, _ in data), sum(s for *_, s in data)
for a, e, s in data:
    ch, _ = off2.get(a - base, ((), "?"))
    key = None
    for f, ln in reversed(ch):
        if f in ("grouped.cuh", "model.cuh", "hmc_cc.cuh", "hmc_mg.cuh"):
            key = (f, ln)
            break
    if key is None and ch:
        key = ch[-1]
    byl[key] += e
    sm[key] += s
    byf[key[0] if key else None] += e
    smf[key[0] if key else None] += s
print("total instructions", tot, "samples", ts)
for k, e in sorted(byf.items(), key=lambda kv: -kv[1]):
    print(f"{str(k):20s} {100 * e / tot:5.1f}% instr {100 * smf[k] / ts:5.1f}% samples")
for k, s in sorted(sm.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{str(k):32s} {100 * byl[k] / tot:5.1f}% instr {100 * s / ts:5.1f}% samples")
