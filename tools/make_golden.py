#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE
(oracle/_ref/libref{64,128}.so, built by oracle/build_ref.sh from /root/reference).

Run here (where /root/reference exists):   python tools/make_golden.py [NAME ...]   (default: all)
The fixtures are committed so that the oracle and the CUDA path can be checked against the
reference's own outputs on the GPU box, where the reference tree does not exist.

Per fixture NAME (parameters in FIXTURES below):
  NAME.lph        index built by the reference's build-p (mphf::build + essentials::save)
  NAME.npz        index_bases/index_offsets   the synthetic unitigs the index was built from
                  q_bases/q_offsets           query batch: the unitigs, mutated/random reads,
                                              reads with N / lower case, too-short and empty contigs
                  q_codes/q_code_offsets      reference hf(contig, len, true) for every query contig
                  rec                         minimizer::from_string stream of the index set (18 B records)
                  n_kmers, mm_count           its return value / final m-mer ordinal
                  triplets, coll_ids          minimizer::classify outputs
                  coll_kmers                  minimizer::get_colliding_kmers stream
                  csv                         the build-p CSV line
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lphash_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

#            name                k   m  bits n_kmers seed
FIXTURES = [("k31_m20_u64", 31, 20, 64, 24000, 0xA001),
            ("k31_m16_u128", 31, 16, 128, 24000, 0xA002),
            ("k63_m24_u128", 63, 24, 128, 24000, 0xA003),
            ("k47_m20_u128", 47, 20, 128, 16000, 0xA004),
            ("k15_m7_u64", 15, 7, 64, 12000, 0xA005),
            ("k21_m11_u64", 21, 11, 64, 12000, 0xA006),
            # (k, m) pairs the tiled kernel is NOT instantiated for: the generic query / scan kernels
            ("k25_m13_u64", 25, 13, 64, 12000, 0xA007),
            ("k40_m17_u128", 40, 17, 128, 12000, 0xA008)]


def make_queries(bases, offsets, k, m, seed):
    """Query batch exercising members, non-members, seams, invalid bytes and degenerate lengths."""
    rng = np.random.Generator(np.random.PCG64(seed))
    genome = bases[: int(offsets[min(len(offsets) - 1, 8)])]  # a prefix of the unitigs as "genome"
    recs = []
    raw = bases.tobytes()
    for c in range(len(offsets) - 1):  # 1. the index set itself (all members)
        recs.append(raw[int(offsets[c]):int(offsets[c + 1])])
    rb, ro = synth.reads(160, genome, read_len=150, seed=seed + 1)  # 2. mutated / random reads
    rraw = rb.tobytes()
    for c in range(len(ro) - 1):
        recs.append(rraw[int(ro[c]):int(ro[c + 1])])
    # 3. degenerate lengths: empty, < m, == m, < k, == k, k + 1
    for ln in (0, 1, m - 1, m, k - 1, k, k + 1, 2 * k):
        recs.append(synth.random_bases(ln, rng).tobytes())
    # 4. lower / mixed case and U (all valid per the reference's table)
    s = bytearray(recs[0][: 3 * k])
    recs.append(bytes(s).lower())
    recs.append(bytes(s).replace(b"T", b"U"))
    recs.append(bytes(s).replace(b"T", b"u").replace(b"A", b"a"))
    # 5. invalid bytes (the reference's streaming quirk): single N, runs of N, N near the ends,
    #    N closer than k / than m to each other, other junk bytes
    order = np.argsort(-np.diff(offsets).astype(np.int64), kind="stable")[:6]  # the longest unitigs
    dirty_src = [raw[int(offsets[c]):int(offsets[c + 1])] for c in order]
    for j, src in enumerate(dirty_src):
        s = bytearray(src[: 40 * k])
        n = len(s)
        pos = sorted(set(int(x) for x in rng.integers(0, n, size=3 + 2 * j)))
        for p in pos:
            s[p] = ord("N")
        recs.append(bytes(s))
    s = bytearray(dirty_src[0][: 10 * k])
    for p in (0, 1, m, k - 1, k, 2 * k + 3, 2 * k + 3 + m, 2 * k + 4 + 2 * m, len(s) - 1, len(s) - m, len(s) - k):
        s[p % len(s)] = ord("n")
    recs.append(bytes(s))
    s = bytearray(dirty_src[1][: 12 * k])
    n = len(s)
    s[(3 * k) % (n - 5):(3 * k) % (n - 5) + 5] = b"NNNNN"
    s[(5 * k) % n] = ord("-")
    s[(5 * k + 2) % n] = ord("R")
    s[(7 * k) % n] = 0
    s[(8 * k + 1) % n] = 255
    recs.append(bytes(s))
    for c in range(24):  # short reads with sparse N, like FASTQ
        lo = int(rng.integers(0, max(1, len(genome) - 151)))
        s = bytearray(genome[lo:lo + 150].tobytes())
        for p in rng.integers(0, 150, size=int(rng.integers(1, 4))):
            s[int(p)] = ord("N")
        recs.append(bytes(s))
    recs.append(b"N" * (2 * k))
    recs.append(b"")
    lens = np.array([len(r) for r in recs], dtype=np.uint64)
    q_offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=q_offsets[1:])
    q_bases = np.frombuffer(b"".join(recs), dtype=np.uint8).copy()
    return recs, q_bases, q_offsets


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, k, m, bits, n_kmers, seed in FIXTURES:
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            bases, offsets = synth.unitigs(n_kmers, k, m, seed=seed, min_len=2 * k, max_len=3000,
                                           planted=8)
            fa = os.path.join(tmp, "index.fa")
            synth.write_fasta(fa, bases, offsets)
            lph = os.path.join(OUT, name + ".lph")
            csv = ref.build(fa, k, m, lph, bits=bits, tmp_dir=tmp)
            f = ref.RefMphf(lph, bits)
            recs, q_bases, q_offsets = make_queries(bases, offsets, k, m, seed)
            codes = [f.query(r) for r in recs]
            q_code_offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
            np.cumsum([len(c) for c in codes], out=q_code_offsets[1:])
            q_codes = np.concatenate(codes) if codes else np.zeros(0, np.uint64)
            rec, nk, mm = ref.scan(bases, offsets, k, m, bits=bits)
            trip, ids = ref.classify(bases, offsets, k, m, bits=bits)
            ck = ref.colliding_kmers(bases, offsets, k, m, ids, bits=bits)
            np.savez_compressed(os.path.join(OUT, name + ".npz"),
                                k=k, m=m, bits=bits, index_bases=bases, index_offsets=offsets,
                                q_bases=q_bases, q_offsets=q_offsets, q_codes=q_codes,
                                q_code_offsets=q_code_offsets, rec=rec, n_kmers=nk, mm_count=mm,
                                triplets=trip, coll_ids=ids, coll_kmers=ck, csv=csv)
            print(f"{name}: {f.kmer_count} k-mers, lph {os.path.getsize(lph)} B, "
                  f"{len(recs)} query contigs -> {len(q_codes)} codes, {len(rec)} records, "
                  f"{len(ids)} colliding ids, {len(ck)} colliding k-mers | {csv}")


if __name__ == "__main__":
    main()
