"""Per-region executed-instruction profile from an `ncu --page source --csv` dump (tools, not product)."""
import csv
import sys


def main(path, per_unit, chunk=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, ie, isamp, iat = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples",
                                                        "Avg. Threads Executed"))
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), r[isrc], int(r[ie]), int(r[isamp]),
                         float(r[iat] or 0)))
        except (ValueError, IndexError):
            pass
    total = sum(d[2] for d in data)
    tsamp = sum(d[3] for d in data)
    print("total exec", total, "per unit", total / per_unit, "samples", tsamp)
    base = data[0][0]
    for i in range(0, len(data), chunk):
        seg = data[i:i + chunk]
        ex = sum(d[2] for d in seg)
        if ex / total > 0.004:
            print(f"{i:5d} addr {seg[0][0] - base:6x} exec/unit {ex / per_unit:7.1f} share {100 * ex / total:5.1f}% "
                  f"samples {100 * sum(d[3] for d in seg) / tsamp:5.1f}% avgthr {sum(d[4] * d[2] for d in seg) / max(ex, 1):5.1f}  "
                  f"{seg[0][1][:50]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
