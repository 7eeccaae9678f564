#!/usr/bin/env bash
# ncu full capture (steady state, no cache flush) of the query kernel.  Usage: tools/gpu_prof2.sh tag label...
set -u
TAG="$1"; shift
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
for impl in "$@"; do
  timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on \
    -k regex:k_query_ -s 8 -c 1 -f -o "$OUT/prof_$impl" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/prof_$impl.log" 2>&1
  echo "prof $impl exit $?"
done
ls -la "$OUT"
