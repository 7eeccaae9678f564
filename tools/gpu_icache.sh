#!/usr/bin/env bash
# instruction-cache counters of the query kernel for the given library variants.  Usage: tools/gpu_icache.sh tag variant...
set -u
TAG="$1"; shift
OUT="gpurun_out/$TAG"; mkdir -p "$OUT"
M=gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for v in "$@"; do
  lib="$PWD/lphash_b200/liblphash_b200_$v.so"; [ "$v" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  for impl in tiled; do
    LPHB_BENCH_NOCHECK=1 LPHASH_B200_LIB="$lib" timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:k_query_ -s 5 -c 1 --csv --log-file "$OUT/ic_${v}_$impl.csv" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/ic_${v}_$impl.log" 2>&1
    echo "== $v $impl"; python - "$OUT/ic_${v}_$impl.csv" <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
for r in rows[1:]:
    print(f"   {r[-3]:80s} {r[-1]}")
PY
  done
done
