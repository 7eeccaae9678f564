#!/usr/bin/env python
"""Per stall reason: total samples and the top source lines.  Usage: tools/ncu_stalls.py rep [reason...]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r[0] == "Line No": hdr = r
    elif hdr and r[0].isdigit():
        agg.append((cur, int(r[0]), r[1].strip(), r))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def num(r, n):
    v = r[hdr.index(n) - len(hdr)]  # from the right: source text may contain unescaped commas
    return int(v) if v.lstrip('-').isdigit() else 0
tot = {n: sum(num(a[3], n) for a in agg) for n in reasons}
all_ = sum(tot.values())
print("stall samples by reason:", ", ".join(f"{n[6:]}={100*v/all_:.1f}%" for n, v in sorted(tot.items(), key=lambda x: -x[1]) if v))
for n in sys.argv[2:]:
    n = "stall_" + n
    print("==", n)
    for a in sorted(agg, key=lambda a: -num(a[3], n))[:14]:
        print(f"  {a[0]:18s} {a[1]:4d} {100*num(a[3], n)/all_:5.2f}%  {a[2][:100]}")
