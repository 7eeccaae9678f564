#!/usr/bin/env bash
# occupancy sweep (experiment): bench with extra dynamic smem per CTA.  Usage: tools/gpu_occ.sh tag impl pad...
set -u
TAG="$1"; IMPL="$2"; shift 2
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
for pad in "$@"; do
  LPHB_QUERY_IMPL=$IMPL LPHB_EXP_SMEM_PAD=$pad timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_${IMPL}_$pad.json" 2> "$OUT/bench_${IMPL}_$pad.err"
  python - "$IMPL pad=$pad" "$OUT/bench_${IMPL}_$pad.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:22s} kernel_ms {d['roofline']['kernel_ms']:.4f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
