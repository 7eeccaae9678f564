#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference (sources read where they lie under
# $LPHASH_REF_DIR, default /root/reference) into oracle/_ref/:
#   libref64.so  / libref128.so   reference + oracle/ref_harness.cpp (extern "C" entry points)
#   lphash64     / lphash128      the reference's own CLI (src/lphash.cpp), both kmer_t flavours
# The reference selects kmer_t at compile time through include/compile_constants.tpd, which is
# included with a quoted (same-directory) include, so — exactly like the reference's own
# scripts/experiments.sh:76 — each flavour is compiled from a throw-away copy of include/ + src/
# under $TMPDIR with that one-line file swapped.  Nothing is copied into this repository;
# outputs go only to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
# GCC 13 needs `-include memory` (essentials.hpp:371 uses std::unique_ptr without <memory>).
# -march=x86-64-v3 (+bmi2/sse4.2 as in the reference's CMakeLists.txt:23-26) instead of
# -march=native because the binaries are built here and run on the GPU box's host CPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${LPHASH_REF_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/include" ]; then
  echo "build_ref.sh: reference tree not found at $REF (ok on the GPU box: prebuilt _ref is used)" >&2
  exit 0
fi
mkdir -p "$OUT"
# -DNDEBUG: the reference builds as CMake "Release" by default (CMakeLists.txt:5-7), i.e. with its asserts off
CXXFLAGS="-std=c++17 -O3 -DNDEBUG -march=x86-64-v3 -mbmi2 -msse4.2 -pthread -include memory -w -fPIC"
SRCS="src/constants.cpp src/quartet_wtree.cpp src/minimizer.cpp src/partitioned_mphf.cpp src/unpartitioned_mphf.cpp src/mphf_utils.cpp"
build_flavour() {
  local bits="$1" type="$2"
  local tmp; tmp="$(mktemp -d)"
  trap 'rm -rf "$tmp"' RETURN
  cp -r "$REF/include" "$REF/src" "$tmp/"
  chmod -R u+w "$tmp"
  ln -s "$REF/external" "$tmp/external"
  echo "typedef $type kmer_t;" > "$tmp/include/compile_constants.tpd"
  ( cd "$tmp" && g++ $CXXFLAGS -shared -I"$tmp" -o "$OUT/libref$bits.so" "$HERE/ref_harness.cpp" $SRCS -lz ) &
  ( cd "$tmp" && g++ $CXXFLAGS -o "$OUT/lphash$bits" src/lphash.cpp $SRCS src/parser_build.cpp -lz ) &
  wait
}
build_flavour 64 uint64_t
build_flavour 128 __uint128_t
ls -la "$OUT"
