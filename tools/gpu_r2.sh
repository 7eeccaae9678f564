#!/usr/bin/env bash
# Round-2 GPU call: parity tests, then kernel time per library variant, then one ncu capture.
# Usage: tools/gpu_r2.sh <tag> "<variants...>" [prof_variant|none] [notests]
set -u
TAG="$1"; VARS="${2:-default}"; PROF="${3:-default}"; NOTESTS="${4:-}"
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > "$OUT/gpu.csv" 2>&1
if [ -z "$NOTESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -15 "$OUT/pytest_gpu.log"
fi
for v in $VARS; do
  lib="$PWD/lphash_b200/liblphash_b200_$v.so"; [ "$v" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  LPHASH_B200_LIB="$lib" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
  python - "$v" "$OUT/bench_$v.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:12s} kernel_ms {d['roofline']['kernel_ms']:.4f}  ms/step {d['ms_per_step']:.4f}  value {d['value']:.4g}  frac {d['roofline']['frac']:.3f} e2e {d['e2e']['value']:.4g}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 "$OUT/bench_$v.err"
done
if [ "$PROF" != none ]; then
  lib="$PWD/lphash_b200/liblphash_b200_$PROF.so"; [ "$PROF" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  LPHASH_B200_LIB="$lib" timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on \
    -k regex:k_query_tiled -s 8 -c 1 -f -o "$OUT/prof_$PROF" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/prof_$PROF.log" 2>&1
  echo "prof exit $?"; ls -la "$OUT" | tail -5
fi
