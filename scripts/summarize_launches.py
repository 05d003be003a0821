#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel count, total and mean duration, share of the listed GPU time.

    python scripts/summarize_launches.py gpurun_out/<tag>/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import sys


def main():
    rows = []
    with open(sys.argv[1]) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rd:
        if len(r) <= iv:
            continue
        val = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        name = r[ik].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
    total = sum(a[1] for a in agg.values())
    print(f"# source: {sys.argv[1]}")
    print("# ncu launch list: per-launch times are cold-cache and serialised; compare SHARES, not absolutes")
    print(f"# total listed GPU time {total:.1f} us over {sum(a[0] for a in agg.values())} launches")
    print(f"{'kernel':92s} {'n':>5s} {'total_us':>10s} {'mean_us':>9s} {'min_us':>8s} {'max_us':>8s} {'share':>7s}")
    for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{name:92s} {a[0]:5d} {a[1]:10.1f} {a[1] / a[0]:9.2f} {a[2]:8.2f} {a[3]:8.2f} {100 * a[1] / total:6.1f}%")


if __name__ == "__main__":
    main()
