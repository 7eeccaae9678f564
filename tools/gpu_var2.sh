#!/usr/bin/env bash
# kernel time of library variants x kernel impls x optional env.  Usage: tools/gpu_var2.sh tag variant...
set -u
TAG="$1"; shift
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
for v in "$@"; do
  lib="$PWD/lphash_b200/liblphash_b200_$v.so"; [ "$v" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  for impl in pipe tiled; do
    for w in 0 1; do
      extra=""; [ $w = 1 ] && extra="LPHB_NO_L2_WINDOW=1"
      env $extra LPHB_BENCH_NOCHECK=1 LPHB_QUERY_IMPL=$impl LPHASH_B200_LIB="$lib" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/b_${v}_${impl}_$w.json" 2> "$OUT/b_${v}_${impl}_$w.err"
      python - "$v $impl nowindow=$w" "$OUT/b_${v}_${impl}_$w.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:32s} kernel_ms {d['roofline']['kernel_ms']:.4f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
    done
  done
done
