"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,...` launch list per kernel/grid.
usage: python profiles/summarize.py <launches.csv>"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, mi, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Metric Name", "ID"))
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[r[idi]][r[mi]] = float(r[vi].replace(",", ""))
        per[r[idi]]["k"] = r[ki].split("(")[0][-60:]
        per[r[idi]]["g"] = r[gi]
    agg = collections.defaultdict(list)
    for d in per.values():
        agg[(d["k"], d["g"])].append((d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0),
                                      d.get("dram__bytes_write.sum", 0.0)))
    tot = sum(x[0] for v in agg.values() for x in v)
    print("| kernel | grid | launches | avg us | share of listed time | avg DRAM read MB | avg DRAM write MB | GB/s |")
    print("|---|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
        n = len(v)
        t = sum(x[0] for x in v) / n / 1e3
        rd = sum(x[1] for x in v) / n / 1e6
        wr = sum(x[2] for x in v) / n / 1e6
        print(f"| `{k[0]}` | {k[1]} | {n} | {t:.2f} | {sum(x[0] for x in v) / tot:.1%} | {rd:.2f} | {wr:.3f} | {rd / t * 1e3 if t else 0:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
