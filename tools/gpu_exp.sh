set -u
# kernel time of the fused build scan per library variant: tools/gpu_exp.sh "<variants>"
for v in $1; do
  lib="$PWD/lphash_b200/liblphash_b200_$v.so"; [ "$v" = default ] && lib="$PWD/lphash_b200/liblphash_b200.so"
  LPHASH_B200_LIB="$lib" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cfg5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'scan kernel_ms', round(d['build_scan']['kernel_ms'],4), 'query kernel_ms', round(d['roofline']['kernel_ms'],4))"
done
