#!/usr/bin/env bash
# ncu full capture of the query kernel in steady state (no cache flush, a late launch).
# Usage: tools/gpu_prof.sh <tag> [variant]
set -u
TAG="$1"; V="${2:-default}"
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
lib="$PWD/lphash_b200/liblphash_b200_$V.so"; [ "$V" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
LPHASH_B200_LIB="$lib" timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on \
  -k regex:k_query_ -s 8 -c 1 -f -o "$OUT/prof_$V" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/prof_$V.log" 2>&1
echo "prof exit $?"; ls -la "$OUT"
