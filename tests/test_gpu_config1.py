"""BASELINE config 1 on the GPU: the reference's bundled data (k=31 m=16, 128-bit kmer_t) through the
C ABI (lphb_query_stream, lphb_scan_superkmers) against what the unmodified reference returned for the
same files (tests/golden/config1/expected.json, made by tools/make_config1.py).  Bit-exact, including
the reference's non-ACGT streaming quirk on ecoli1.fasta (one 4.9 Mbase record with 50 runs of N:
4,896,104 codes, 73 of them spurious) and on the FASTQ (105 of 10,000 reads contain N)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from lphash_b200 import api, seqio

pytestmark = pytest.mark.gpu
CFG1 = os.path.join(GOLDEN_DIR, "config1")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def index():
    f = api.Mphf.load(os.path.join(CFG1, "se.ust.k31_m16_u128.lph"), 128)
    yield f
    f.close()


@pytest.mark.parametrize("name", ["self", "salmonella", "ecoli1", "srr"])
def test_bundled_query_files_match_reference(index, name):
    exp = json.load(open(os.path.join(CFG1, "expected.json")))["queries"][name]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, exp["file"]))
    codes, code_off = index.query_batch(bases, offsets)
    assert len(codes) == exp["n_codes"]
    assert [int(x) for x in codes[:8]] == exp["first_codes"]
    assert sha(codes) == exp["sha256_codes"]
    assert sha(np.diff(code_off)) == exp["sha256_counts"]
    if name == "self":  # all members: a minimal perfect hash
        assert np.array_equal(np.sort(codes), np.arange(index.get_kmer_count(), dtype=np.uint64))


def test_bundled_index_scan_matches_reference():
    exp = json.load(open(os.path.join(CFG1, "expected.json")))["scan"]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, "se.ust.k31.fa.gz"))
    rec, nk, mm = api.scan_superkmers(bases, offsets, 31, 16)
    assert (len(rec), nk, mm) == (exp["records"], exp["n_kmers"], exp["mm_count"])
    assert hashlib.sha256(rec.tobytes()).hexdigest() == exp["sha256_records"]


@pytest.mark.parametrize("name", ["ecoli1", "srr"])
def test_build_side_on_bundled_files_with_non_acgt_bytes(name):
    """Parts 1, 2 and 4 of build-p over ecoli1.fasta (50 runs of N) and the FASTQ (105 reads with N) against what
    the reference's from_string / classify / get_colliding_kmers returned for them (include/minimizer.hpp:138-151)"""
    exp = json.load(open(os.path.join(CFG1, "expected.json")))
    e = exp["scan_dirty"][name]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, exp["queries"][name]["file"]))
    rec, nk, mm = api.scan_superkmers(bases, offsets, 31, 16)
    assert (len(rec), nk, mm) == (e["records"], e["n_kmers"], e["mm_count"])
    assert hashlib.sha256(rec.tobytes()).hexdigest() == e["sha256_records"]
    trip, ids, nk2, mm2 = api.scan_classify(bases, offsets, 31, 16)
    assert (nk2, mm2) == (nk, mm)
    assert hashlib.sha256(trip.tobytes()).hexdigest() == e["sha256_triplets"] and sha(ids) == e["sha256_ids"]
    km = api.colliding_kmers(bases, offsets, 31, 16, ids, kmer_bits=128)
    assert len(km) == e["colliding_kmers"]
    assert hashlib.sha256(np.ascontiguousarray(km).tobytes()).hexdigest() == e["sha256_colliding_kmers"]
