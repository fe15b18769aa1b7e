"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share."""
import collections
import csv
import sys


def main(path, last=0):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v, u = float(r[vi].replace(",", "")), r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        seq.append((r[ki].split("(")[0].replace("void ", ""), v))
    if last:
        seq = seq[-last:]
    agg = collections.OrderedDict()
    for n, v in seq:
        agg.setdefault(n, []).append(v)
    tot = sum(v for _, v in seq)
    print(f"{len(seq)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache ncu timings)")
    for n, vs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {n:45s} x{len(vs):4d}  mean {sum(vs) / len(vs):10.1f} us  share {100 * sum(vs) / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
