"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v * 1e3 if u == "s" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    rows = [f"{'total ms':>10} {'launches':>8} {'avg ms':>9} {'share':>6}  kernel"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rows.append(f"{a[1]:10.2f} {a[0]:8d} {a[1] / a[0]:9.3f} {100 * a[1] / tot:5.1f}%  {k[:110]}")
    rows.append(f"total {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches")
    text = "\n".join(rows)
    print(text)
    if out:
        open(out, "a").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
