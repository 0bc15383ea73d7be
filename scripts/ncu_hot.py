"""Print the hottest SASS instructions (by warp-stall samples) of the first kernel in an ncu source-page CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) >= len(hdr): data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("instructions", len(data), "total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[idx[c]])
print("stall totals:", sorted(((v, k[6:]) for k, v in agg.items()), reverse=True)[:10])
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[idx[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(i, r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[idx["Source"]].strip()[:80], st)
