"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (argument: the csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: 0 for h in stalls}
for s, r in data:
    for h in stalls:
        try:
            agg[h] += int(r[ix[h]])
        except ValueError:
            pass
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:7d} {100.0 * s / tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:90]:90s} {top}")
