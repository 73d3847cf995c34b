"""Top stall sites of one kernel from `ncu --page source --csv` output (SASS view)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    print(f"{i:5d} {int(r[ix['# Samples']]):6d}  {r[ix['Source']].strip()[:70]:70s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
