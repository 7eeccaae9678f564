#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source-line range (code footprint: the I-cache holds
32 KB = 2048 instructions per SM).  Usage: tools/sass_footprint.py obj.o kernel_substr [file:lo-hi:label ...]"""
import re, subprocess, sys, tempfile, os, glob
obj, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = glob.glob(d + "/*.cubin")[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = None; counts = {}; total = 0
for ln in txt:
    if ln.startswith(".text."):
        inside = kern in ln
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        total += 1
        counts[cur] = counts.get(cur, 0) + 1
print(f"{kern}: {total} SASS instructions = {total*16/1024:.1f} KB")
rest = total
for spec in sys.argv[3:]:
    f, rng, label = (spec.split(":") + [spec])[:3]
    lo, hi = map(int, rng.split("-"))
    n = sum(v for (ff, l), v in counts.items() if ff == f and lo <= l <= hi)
    rest -= n
    print(f"  {label:24s} {n:6d}")
if len(sys.argv) > 3: print(f"  {'(other)':24s} {rest:6d}")
else:
    byf = {}
    for (ff, l), v in counts.items(): byf[ff] = byf.get(ff, 0) + v
    for k, v in sorted(byf.items(), key=lambda x: -x[1]): print(f"  {k:28s} {v}")
