#!/usr/bin/env bash
# Bench several tuning variants of the library in one gpurun call.  Usage: tools/gpu_variants.sh <tag> name...
# ("default" = lphash_b200/liblphash_b200.so; others are built by `make -C lphash_b200/csrc variant`).
# Prints kernel_ms / value per variant.
set -u
TAG="$1"; shift
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
for v in "$@"; do
  lib="$PWD/lphash_b200/liblphash_b200_$v.so"; [ "$v" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  LPHASH_B200_LIB="$lib" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
  python - "$v" "$OUT/bench_$v.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:12s} kernel_ms {d['roofline']['kernel_ms']:.4f}  ms/step {d['ms_per_step']:.4f}  value {d['value']:.4g}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
