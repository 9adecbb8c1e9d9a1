"""Executed warp-instructions per phase of the sampler kernel (tools, not product).

Joins `nvdisasm -gi` (inline chains) with an `ncu --page source --csv` dump and attributes every SASS instruction
to the outermost-but-one frame that lies inside dm_grad / dm_value (model.cuh) or the kernel body
(kernels_impl.cuh), so that inlined special functions count for the phase that called them.

usage: ncu_phases.py <nvdisasm -gi dump> <mangled kernel> <source csv> <units per launch> [phase table]
"""
import collections
import csv
import re
import sys

# (file, first line, last line, label) — edit when the sources move
PHASES = [
    ("hmc_cc.cuh", 312, 316, "cc: effects"),
    ("hmc_cc.cuh", 317, 329, "cc: concentrations, totals, pair records"),
    ("hmc_cc.cuh", 330, 341, "cc: rational + items + odd counts"),
    ("hmc_cc.cuh", 342, 355, "cc: chain rule"),
    ("hmc_cc.cuh", 359, 438, "cc: value"),
    ("hmc_cc.cuh", 443, 510, "cc: kernel prologue"),
    ("hmc_cc.cuh", 511, 526, "cc: momentum draw + first kick"),
    ("hmc_cc.cuh", 527, 542, "cc: leapfrog updates"),
    ("hmc_cc.cuh", 543, 564, "cc: value call + accept"),
    ("hmc_cc.cuh", 565, 585, "cc: adapt + trace"),
    ("grouped.cuh", 140, 177, "P0 effects (grouped)"),
    ("grouped.cuh", 178, 224, "P1 pair records + totals (grouped)"),
    ("grouped.cuh", 225, 247, "E items (grouped)"),
    ("grouped.cuh", 248, 330, "C columns + chain rule + epilogue (grouped)"),
    ("model.cuh", 167, 203, "F1 effects"),
    ("model.cuh", 204, 266, "F2/F3 concentrations, row terms"),
    ("model.cuh", 267, 323, "E elements"),
    ("model.cuh", 324, 478, "B column sums + chain rule + epilogue"),
    ("model.cuh", 479, 640, "value"),
    ("kernels_impl.cuh", 0, 100000, "sampler (momentum, accept, adapt, trace)"),
]

PIPE = [
    (r"^(FFMA|FMUL|FADD|HFMA2|IMAD|FMNMX3?)", "fma"),
    (r"^(MUFU)", "xu"),
    (r"^(LDS|STS|LDG|STG|LDC|LDCU|ATOMS|LD|ST)\b", "lsu"),
    (r"^(D[A-Z]+)", "fp64"),
    (r"^(SHFL|BAR|WARPSYNC|VOTE|BRA|BSSY|BSYNC|CALL|RET|EXIT|NANOSLEEP)", "ctl/shfl"),
]


def classify(op):
    for pat, name in PIPE:
        if re.match(pat, op):
            return name
    return "alu"


def main(disasm, func, srccsv, per_unit):
    lines = open(disasm).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section") and (".text." + func) in l)
    chain = []
    off2 = {}
    pending_new = True
    for l in lines[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            if pending_new:
                chain = []
                pending_new = False
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?(\S+)", l)
        if m:
            off2[int(m.group(1), 16)] = (tuple(chain), m.group(2).rstrip(";"))
            pending_new = True
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
    by_phase = collections.Counter()
    by_phase_pipe = collections.defaultdict(collections.Counter)
    by_line = collections.Counter()
    samp_phase = collections.Counter()
    for a, e, s in data:
        ch, op = off2.get(a - base, ((), "?"))
        label, key = "other", None
        for f, ln in ch:
            hit = next((p for p in PHASES if p[0] == f and p[1] <= ln <= p[2]), None)
            if hit:
                label, key = hit[3], (f, ln)
                break
        by_phase[label] += e
        samp_phase[label] += s
        by_phase_pipe[label][classify(op)] += e
        by_line[key] += e
    tot = sum(by_phase.values())
    ts = sum(samp_phase.values())
    print(f"total warp-instructions per unit: {tot / per_unit:.1f}")
    for p in [p[3] for p in PHASES] + ["other"]:
        e = by_phase[p]
        pipes = "  ".join(f"{k} {v / per_unit:6.1f}" for k, v in sorted(by_phase_pipe[p].items()))
        print(f"{p:45s} {e / per_unit:7.1f} ({100 * e / tot:4.1f}%, samples {100 * samp_phase[p] / max(ts, 1):4.1f}%)   {pipes}")
    if len(sys.argv) > 5:
        for k, e in sorted(by_line.items(), key=lambda kv: (kv[0] is None, kv[0])):
            if e / tot > 0.004:
                print(f"   {str(k):32s} {e / per_unit:7.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]))
