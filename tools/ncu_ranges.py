#!/usr/bin/env python
"""Per source-line-range instruction / stall-sample shares of an ncu report.
Usage: tools/ncu_ranges.py report.ncu-rep file:lo-hi[:label] ...   (n_kmers via env NK)"""
import csv, os, subprocess, sys

def main():
    rep = sys.argv[1]
    nk = float(os.environ.get("NK", "100004736"))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]
        elif r[0] == "Line No": hdr = r
        elif hdr and r[0].isdigit():
            def num(name):
                v = r[hdr.index(name) - len(hdr)]
                return int(v) if v.lstrip("-").isdigit() else 0
            agg.append((cur, int(r[0]), num("Instructions Executed"), num("# Samples")))
    tot = sum(a[2] for a in agg) or 1
    tots = sum(a[3] for a in agg) or 1
    print(f"total warp inst {tot}  = {tot*32/nk:.1f} thread-inst/k-mer")
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        f, rng = parts[0], parts[1]
        label = parts[2] if len(parts) > 2 else spec
        lo, hi = (int(x) for x in rng.split("-"))
        i = sum(a[2] for a in agg if a[0] == f and lo <= a[1] <= hi)
        s = sum(a[3] for a in agg if a[0] == f and lo <= a[1] <= hi)
        print(f"{label:28s} inst {100*i/tot:5.1f}% ({i*32/nk:6.1f}/k-mer)  samples {100*s/tots:5.1f}%")
main()
