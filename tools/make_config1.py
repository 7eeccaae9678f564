#!/usr/bin/env python
"""BASELINE config 1 as committed fixtures: the reference's bundled data (se.ust.k31 unitigs, k=31 m=16,
default __uint128_t kmer_t, seed 42, c 3.0) built and queried by THE UNMODIFIED REFERENCE
(oracle/_ref, built by oracle/build_ref.sh from /root/reference), so that the oracle and the CUDA path
can be checked against it where the reference tree does not exist.

Run here:   python tools/make_config1.py
Writes tests/golden/config1/
  se.ust.k31.fa.gz, salmonella_enterica.fasta.gz, ecoli1.fasta.gz, SRR5833294.10K.fastq.gz
                      byte copies of the reference's data files (inputs only; not source code)
  se.ust.k31_m16_u128.lph   index built by the reference's build-p
  expected.json       per query file: number of records, number of codes, 64-bit FNV fold of SURVEY 8c,
                      sha256 of the little-endian u64 codes, sha256 of the per-record code counts, first codes;
                      from_string totals of the index set
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lphash_b200 import seqio  # noqa: E402
from oracle import ref  # noqa: E402

REF_DATA = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden", "config1")
K, M, BITS = 31, 16, 128
FILES = {"self": "unitigs_stitched/se.ust.k31.fa.gz", "salmonella": "queries/salmonella_enterica.fasta.gz",
         "ecoli1": "queries/ecoli1.fasta.gz", "srr": "queries/SRR5833294.10K.fastq.gz"}


def fnv(codes: np.ndarray) -> str:
    h = 0xCBF29CE484222325
    for v in codes.tolist():
        h = ((h ^ v) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def main():
    os.makedirs(OUT, exist_ok=True)
    for rel in FILES.values():
        shutil.copyfile(os.path.join(REF_DATA, rel), os.path.join(OUT, os.path.basename(rel)))
    lph = os.path.join(OUT, "se.ust.k31_m16_u128.lph")
    with tempfile.TemporaryDirectory() as tmp:
        csv = ref.build(os.path.join(REF_DATA, FILES["self"]), K, M, lph, bits=BITS, tmp_dir=tmp)
    f = ref.RefMphf(lph, BITS)
    exp = {"k": K, "m": M, "kmer_bits": BITS, "build_csv": csv, "nkmers": int(f.kmer_count), "queries": {}}
    for name, rel in FILES.items():
        bases, offsets = seqio.read_batch(os.path.join(REF_DATA, rel))
        raw = bases.tobytes()
        # one hf(seq, len, true) per record, like the reference driver (src/query.cpp:51-52); records
        # with non-ACGT bytes take the reference's streaming quirk
        per = [f.query(raw[int(offsets[c]):int(offsets[c + 1])]) for c in range(len(offsets) - 1)]
        codes = np.concatenate(per) if per else np.zeros(0, np.uint64)
        counts = np.array([len(x) for x in per], dtype=np.uint64)
        exp["queries"][name] = {
            "file": os.path.basename(rel), "records": int(len(offsets) - 1), "bases": int(offsets[-1]),
            "n_codes": int(len(codes)), "fnv": fnv(codes),
            "sha256_codes": hashlib.sha256(np.ascontiguousarray(codes, dtype="<u8").tobytes()).hexdigest(),
            "sha256_counts": hashlib.sha256(np.ascontiguousarray(counts, dtype="<u8").tobytes()).hexdigest(),
            "first_codes": [int(x) for x in codes[:8]]}
        print(name, exp["queries"][name])
    bases, offsets = seqio.read_batch(os.path.join(REF_DATA, FILES["self"]))
    rec, nk, mm = ref.scan(bases, offsets, K, M, bits=BITS)
    exp["scan"] = {"records": int(len(rec)), "n_kmers": int(nk), "mm_count": int(mm),
                   "sha256_records": hashlib.sha256(rec.tobytes()).hexdigest()}
    print("scan", exp["scan"])
    # the reference's build-side loops on input WITH non-ACGT bytes (include/minimizer.hpp:138-151, 283-300):
    # ecoli1.fasta (50 runs of N) and the FASTQ (105 reads with N) through from_string, classify and
    # get_colliding_kmers
    exp["scan_dirty"] = {}
    for name in ("ecoli1", "srr"):
        bases, offsets = seqio.read_batch(os.path.join(REF_DATA, FILES[name]))
        rec, nk, mm = ref.scan(bases, offsets, K, M, bits=BITS)
        trip, ids = ref.classify(bases, offsets, K, M, bits=BITS)
        km = ref.colliding_kmers(bases, offsets, K, M, ids, bits=BITS)
        exp["scan_dirty"][name] = {"records": int(len(rec)), "n_kmers": int(nk), "mm_count": int(mm),
                                   "sha256_records": hashlib.sha256(rec.tobytes()).hexdigest(),
                                   "triplets": int(len(trip)), "colliding_ids": int(len(ids)),
                                   "sha256_triplets": hashlib.sha256(trip.tobytes()).hexdigest(),
                                   "sha256_ids": hashlib.sha256(np.ascontiguousarray(ids, dtype="<u8").tobytes()).hexdigest(),
                                   "colliding_kmers": int(len(km)),
                                   "sha256_colliding_kmers": hashlib.sha256(np.ascontiguousarray(km).tobytes()).hexdigest()}
        print("scan_dirty", name, exp["scan_dirty"][name])
    json.dump(exp, open(os.path.join(OUT, "expected.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
