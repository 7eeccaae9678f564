#!/usr/bin/env python
"""BASELINE config 3 input preparation (CPU only, ~20-40 min, ~20 GB of scratch): the human-scale
synthetic unitig set (2.5e9 k-mers, k=31 m=20, recipe of config 2) and its index, built by the
reference's own build-p with c=5.0 (the paper's human setting, scripts/experiments.sh:71,131).
Writes bench_cache/cfg3_n2500000000_k31_m20_u64.lph (git-ignored; travels to the GPU box with gpurun).
Usage: python tools/build_cfg3_index.py [n_kmers]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lphash_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000_000
# A random genome of 2.5e9 bases holds a repeated 31-mer with probability ~0.5 (n^2 / (2 * 4^31) = 0.68
# expected pairs); the reference's build then fails in PTHash ("seed did not work": the repeated k-mer is a
# duplicate key of fallback_kmer_order).  The seed is therefore part of the workload definition:
# tools/build_cfg3_index.py <n> <seed> tries one, bench_cache/cfg3_seed.txt records the one that built.
seed = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x5EED0013
cache = os.path.join(ROOT, "bench_cache")
os.makedirs(cache, exist_ok=True)
tag = f"cfg3_n{n}_k31_m20_u64"
lph = os.path.join(cache, tag + ".lph")
t0 = time.time()
bases, offsets = synth.unitigs(n, 31, 20, seed=seed)
print(f"unitigs: {len(offsets) - 1} contigs, {len(bases)} bases ({time.time() - t0:.0f}s)", flush=True)
fa = os.path.join(cache, tag + ".fa")
synth.write_fasta(fa, bases, offsets)
del bases
print(f"fasta written ({time.time() - t0:.0f}s)", flush=True)
csv = ref.build(fa, 31, 20, lph + ".tmp", bits=64, c=5.0, threads=8, max_memory_gb=24, tmp_dir=cache)
os.replace(lph + ".tmp", lph)
os.remove(fa)
open(os.path.join(cache, "cfg3_seed.txt"), "w").write(hex(seed) + "\n")
print(f"build-p (genome seed {seed:#x}): {csv} ({time.time() - t0:.0f}s), {os.path.getsize(lph)} bytes", flush=True)
