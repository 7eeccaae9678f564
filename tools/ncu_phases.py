#!/usr/bin/env python
"""Executed instructions and stall samples of k_query_tiled per phase (A pack, B scan, C list, D probe,
E emit, cold paths), from an ncu report with source info.  Phases are found from the `// ---- X:` markers
and function heads of the query_tiled.cu the report was built from (pass its path as 2nd argument).
Usage: tools/ncu_phases.py report.ncu-rep [query_tiled.cu] [n_kmers]"""
import csv
import re
import subprocess
import sys


def ranges(src):
    """(first_line, label) sorted; a line belongs to the last range starting at or before it"""
    out = []
    fn = re.compile(r"^(?:static\s+)?(?:__device__|__global__).*?\b(\w+)\s*\(")
    lines = open(src).read().split("\n")
    for i, l in enumerate(lines, 1):
        m = re.search(r"// -{20,} (\w): ", l)
        if m:
            out.append((i, m.group(1)))
            continue
        m = fn.match(l.strip()) if not l.startswith(" ") else None
        if m:
            name = m.group(1)
            lab = {"codes4": "A", "pack4_top": "A", "bad4": "A", "mark_dirty": "cold", "mul_murmur": "B hash",
                   "mul_murmur_hi": "B hash", "murmur_top32": "B hash", "win16": "B", "window_min": "B min",
                   "mmer_at": "D", "kmer_at": "cold", "exact_strip": "cold", "emit_plain": "E", "emit_masked": "E",
                   "emit_general": "E general", "mark_invalid": "cold", "slow_kmers": "cold",
                   "k_query_tiled": "setup", "k_tile_setup": "setup"}.get(name, "other")
            out.append((i, lab))
    return sorted(out)


def main():
    rep = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else "lphash_b200/csrc/query_tiled.cu"
    nk = float(sys.argv[3]) if len(sys.argv) > 3 else 100004736.0
    rg = ranges(src)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, agg = None, None, {}
    tot_i = tot_s = 0
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
            ln, inst, smp = int(r[0]), num("Instructions Executed"), num("# Samples")
            if cur == "query_tiled.cu":
                lab = "other"
                for a, b in rg:
                    if a <= ln:
                        lab = b
            elif cur == "device_mphf.cuh":
                lab = "D"
            else:
                lab = "intrinsics (" + cur.split(".")[0] + ")"
            a = agg.setdefault(lab, [0, 0])
            a[0] += inst
            a[1] += smp
            tot_i += inst
            tot_s += smp
    print(f"{'phase':34s} {'inst %':>7s} {'inst/k-mer':>11s} {'stall samples %':>16s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{k:34s} {100 * v[0] / tot_i:6.1f}% {v[0] * 32 / nk:11.1f} {100 * v[1] / tot_s:15.1f}%")
    print(f"{'total':34s} {100.0:6.1f}% {tot_i * 32 / nk:11.1f}")


if __name__ == "__main__":
    main()
