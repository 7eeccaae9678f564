#!/usr/bin/env bash
# Literal drop-in build: the reference's OWN driver (src/lphash.cpp, which includes src/query.cpp and
# src/build.cpp) compiled unmodified, with include/partitioned_mphf.hpp shadowed by
# integration/dropin/partitioned_mphf.hpp, linked against lphash_b200/liblphash_b200.so.
# Like oracle/build_ref.sh, the reference's include/ and src/ are copied to a throw-away directory under
# $TMPDIR (the reference selects kmer_t through a same-directory include); nothing of the reference enters
# this repository.  Outputs: oracle/_ref/lphash_gpu64, oracle/_ref/lphash_gpu128 (git-ignored; they travel
# to the GPU box with gpurun, where tests/test_dropin_cli.py runs them beside the reference's CLI).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${LPHASH_REF_DIR:-/root/reference}"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/include" ]; then
  echo "build_dropin.sh: reference tree not found at $REF (ok on the GPU box: prebuilt binaries are used)" >&2
  exit 0
fi
mkdir -p "$OUT"
CXXFLAGS="-std=c++17 -O3 -DNDEBUG -march=x86-64-v3 -mbmi2 -msse4.2 -pthread -include memory -w"
REF_TUS="src/constants.cpp src/quartet_wtree.cpp src/minimizer.cpp src/partitioned_mphf.cpp src/mphf_utils.cpp src/unpartitioned_mphf.cpp src/parser_build.cpp"
build_flavour() {
  local bits="$1" type="$2"
  local tmp; tmp="$(mktemp -d)"
  trap 'rm -rf "$tmp"' RETURN
  cp -r "$REF/include" "$REF/src" "$tmp/"
  chmod -R u+w "$tmp"
  ln -s "$REF/external" "$tmp/external"
  echo "typedef $type kmer_t;" > "$tmp/include/compile_constants.tpd"
  mv "$tmp/include/partitioned_mphf.hpp" "$tmp/include/partitioned_mphf_reference.hpp"   # the original, untouched
  cp "$HERE/partitioned_mphf.hpp" "$tmp/include/partitioned_mphf.hpp"                      # the shadow
  mv "$tmp/include/unpartitioned_mphf.hpp" "$tmp/include/unpartitioned_mphf_reference.hpp"
  cp "$HERE/unpartitioned_mphf.hpp" "$tmp/include/unpartitioned_mphf.hpp"
  cp "$HERE/gpu_build.hpp" "$tmp/include/gpu_build.hpp"
  ( cd "$tmp"
    for f in $REF_TUS; do   # the reference's own translation units: `mphf` is the renamed reference class
      g++ $CXXFLAGS -DLPHASH_B200_REFERENCE_TU -Dmphf=mphf_reference -Dmphf_alt=mphf_alt_reference -I"$ROOT/include" -c "$f" -o "$(basename "$f" .cpp).o" &
    done
    # the driver, unmodified: its `mphf` is the GPU-backed class of the shadow header
    g++ $CXXFLAGS -DLPHASH_B200_KMER_BITS="$bits" -DLPHASH_B200_WITH_ZLIB -I"$ROOT/include" -c src/lphash.cpp -o lphash.o &
    wait
    g++ -o "$OUT/lphash_gpu$bits" lphash.o constants.o quartet_wtree.o minimizer.o partitioned_mphf.o mphf_utils.o \
        unpartitioned_mphf.o parser_build.o -L"$ROOT/lphash_b200" -llphash_b200 -Wl,-rpath,'$ORIGIN/../../lphash_b200' -lz -pthread )
}
build_flavour 64 uint64_t
build_flavour 128 __uint128_t
ls -la "$OUT"/lphash_gpu*
