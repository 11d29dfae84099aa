#!/usr/bin/env python3
"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) -> markdown table of kernels by share of GPU time.
usage: launch_table.py launches.csv "title" > profiles/NAME.md"""
import csv
import sys
from collections import OrderedDict


def main(path, title):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        name = r[ki][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s\n" % title)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (serialised, cold caches: compare SHARES, not absolutes).\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f | %.1f%% |" % (name, n, t, t / n, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "launch list")
