"""Executed instructions per source line: join `nvdisasm -g` line info with an `ncu --page source --csv` dump."""
import collections
import csv
import re
import sys


def main(disasm, func, srccsv, per_unit):
    lines = open(disasm).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + func) or (".text." + func) in l and l.strip().startswith(".section"))
    cur_line = None
    off2line = {}
    for l in lines[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m:
            off2line[int(m.group(1), 16)] = cur_line
    rows = list(csv.reader(open(srccsv)))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), int(r[ie]), int(r[isamp])))
        except (ValueError, IndexError):
            pass
    base = data[0][0]
    ex = collections.Counter()
    sm = collections.Counter()
    for a, e, s in data:
        ln = off2line.get(a - base)
        ex[ln] += e
        sm[ln] += s
    tot = sum(ex.values())
    ts = sum(sm.values())
    for ln, e in sorted(ex.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if e / tot > 0.003:
            print(f"{str(ln):40s} exec/unit {e / per_unit:7.1f}  {100 * e / tot:5.1f}%   samples {100 * sm[ln] / ts:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]))
