"""Per-kernel time and DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list: launches, average time, DRAM bytes per launch, achieved DRAM GB/s and the
fraction of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs, fallback 6650)."""
import collections
import csv
import json
import os
import re
import sys

TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
TO_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 6650.0
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    lines = [l for l in open(path) if l.startswith('"')]
    per_id = collections.OrderedDict()
    for row in csv.DictReader(lines):
        ent = per_id.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]), "ms": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        m, u = row["Metric Name"], row["Metric Unit"]
        if m.startswith("gpu__time_duration"):
            ent["ms"] = v * TO_MS.get(u, 1.0)
        elif m.startswith("dram__bytes_read"):
            ent["rd"] = v * TO_B.get(u, 1.0)
        elif m.startswith("dram__bytes_write"):
            ent["wr"] = v * TO_B.get(u, 1.0)
    agg = collections.OrderedDict()
    for ent in per_id.values():
        a = agg.setdefault(ent["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += ent["ms"]; a[2] += ent["rd"]; a[3] += ent["wr"]
    tot = sum(a[1] for a in agg.values()) or 1.0
    rows = [f"{'total ms':>9} {'launches':>8} {'avg ms':>8} {'share':>6} {'MB rd':>8} {'MB wr':>8} {'GB/s':>7} {'of copy':>7}  kernel"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / (a[1] * 1e-3) / 1e9 if a[1] > 0 else 0.0
        rows.append(f"{a[1]:9.2f} {a[0]:8d} {a[1] / a[0]:8.3f} {100 * a[1] / tot:5.1f}% {a[2] / a[0] / 1e6:8.1f} {a[3] / a[0] / 1e6:8.1f} "
                    f"{gbs:7.0f} {100 * gbs / peak:6.1f}%  {k[:100]}")
    rows.append(f"total {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches; copy bandwidth {peak:.0f} GB/s")
    text = "\n".join(rows)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
