"""Aggregate an `ncu --page source --csv` export into segments between barriers: samples and top stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Source"); isamp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
cur = {"n": 0, "st": {}}; start = 0
for i, r in enumerate(data):
    n = int(r[isamp]); cur["n"] += n
    for c in stall_cols:
        v = int(r[c]) if r[c] else 0
        if v: cur["st"][hdr[c]] = cur["st"].get(hdr[c], 0) + v
    src = r[ia]
    if any(k in src for k in ("BAR.SYNC", "SYNCS", "EXIT", "UBLKCP", "WARPSYNC")) or i == len(data) - 1:
        if cur["n"]:
            top = sorted(cur["st"].items(), key=lambda kv: -kv[1])[:4]
            print(f"{start:5d}-{i:5d} {cur['n']:6d} {100*cur['n']/tot:5.1f}%  {src.strip()[:52]:52s} {top}")
        cur = {"n": 0, "st": {}}; start = i + 1
