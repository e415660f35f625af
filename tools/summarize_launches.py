"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/summarize_launches.py in.csv"""
import csv, collections, re, sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    n = 0
    for row in r:
        if len(row) <= vi or row[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "").replace("ttts::", "")
        t = float(row[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t; n += 1
    tot = sum(a[1] for a in agg.values())
    print("launches %d   total %.3f ms (ncu-serialised, cold-cache: compare SHARES)" % (n, tot / 1e6))
    print("%-64s %7s %12s %7s" % ("kernel", "count", "total us", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %7d %12.1f %6.1f%%" % (k[:64], c, t / 1e3, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
