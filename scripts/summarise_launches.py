"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total device time, launches and share per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"# {path}: total {tot / 1e6:.2f} ms of device time (cold-cache, serialised: compare shares)")
    print(f"{'ms':>10} {'launches':>8} {'avg ms':>9} {'share':>6}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1] / 1e6:10.2f} {a[0]:8d} {a[1] / a[0] / 1e6:9.3f} {100 * a[1] / tot:5.1f}%  {k.split('(')[0][-70:]}")


if __name__ == "__main__":
    main(sys.argv[1])
