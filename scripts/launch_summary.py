"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':66s} {'n':>4s} {'mean us':>9s} {'min':>8s} {'max':>8s} {'share':>6s}")
    for k, v in agg.items():
        print(f"{k:66s} {len(v):4d} {sum(v)/len(v):9.1f} {min(v):8.1f} {max(v):8.1f} {100*sum(v)/tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
