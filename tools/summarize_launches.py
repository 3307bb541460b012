"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("cosmo::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    print("%-58s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "max_ms", "share"))
    for k, (c, v, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-58s %6d %12.3f %10.3f %6.1f%%" % (k[:58], c, v, m, 100 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1])
