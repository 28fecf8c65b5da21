"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept under profiles/."""
import csv, sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1000.0)
        name = r[ki].split("(")[0].replace("void ", "").replace("dgr::", "")[:60]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | avg us | share of captured GPU time |")
    print("|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% |" % (name, n, t / n, 100.0 * t / total))


if __name__ == "__main__":
    main(sys.argv[1])
