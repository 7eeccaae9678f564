#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
executed warp instructions and stall samples.  Usage: tools/ncu_lines.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0] not in ("", "Function Name") and hdr and r[0].isdigit():
            def num(name):
                v = r[hdr.index(name) - len(hdr)]
                return int(v) if v.lstrip("-").isdigit() else 0
            agg.append((cur, int(r[0]), r[1].strip(), num("Instructions Executed"), num("# Samples")))
    tot = sum(a[3] for a in agg) or 1
    tots = sum(a[4] for a in agg) or 1
    print("total warp instructions", tot, "stall samples", tots)
    byfile = {}
    for a in agg:
        f = byfile.setdefault(a[0], [0, 0])
        f[0] += a[3]
        f[1] += a[4]
    for k, v in byfile.items():
        print(f"  {k:20s} inst {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / tots:5.1f}%")
    agg.sort(key=lambda a: -a[3])
    for a in agg[:top]:
        print(f"{a[0]:18s} {a[1]:4d} inst {100 * a[3] / tot:5.1f}% samp {100 * a[4] / tots:5.1f}%  {a[2][:100]}")


if __name__ == "__main__":
    main()
