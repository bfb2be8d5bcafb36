"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+ grid size)."""
import collections
import csv
import re
import sys


def main(path, by_grid=True):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("wsb::", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = {"us": v / 1e3, "ns": v / 1e6, "ms": v, "s": v * 1e3}.get(unit, v)
        key = (name, r[idx["Grid Size"]].replace(" ", "")) if by_grid else (name, "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-44s %-16s %6s %10s %9s %6s" % ("kernel", "grid", "n", "total ms", "avg us", "share"))
    for (k, g), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %-16s %6d %10.3f %9.2f %5.1f%%" % (k[:44], g, c, ms, 1000 * ms / c, 100 * ms / tot))
    print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], by_grid="--no-grid" not in sys.argv)
