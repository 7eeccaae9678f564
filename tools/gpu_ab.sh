#!/usr/bin/env bash
# One gpurun call: GPU parity tests, then one short bench line per label (labels only name the output files).
# Usage: tools/gpu_ab.sh <tag> [label...]
set -u
TAG="$1"; shift
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -5 "$OUT/pytest_gpu.log"
for v in "$@"; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
  python - "$v" "$OUT/bench_$v.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:12s} kernel_ms {d['roofline']['kernel_ms']:.4f}  ms/step {d['ms_per_step']:.4f}  value {d['value']:.4g}  frac {d['roofline']['frac']:.3f} e2e {d['e2e']['value']:.4g}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -3 "$OUT/bench_$v.err"
done
