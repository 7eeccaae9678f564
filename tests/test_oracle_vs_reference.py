"""BASELINE config 1 (the reference's bundled data, k=31 m=16, default 128-bit kmer_t): the CPU oracle
against outputs of THE UNMODIFIED REFERENCE on the same files.

tests/golden/config1/ holds byte copies of the reference's data files, the index its build-p made from
se.ust.k31.fa.gz, and expected.json = what its own streaming query / from_string returned (generated
here by tools/make_config1.py from oracle/_ref; the reference tree does not exist on the GPU box).
The 64-bit FNV folds are also the ones SURVEY.md section 8c lists, hard-coded below as a second anchor.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, fnv_fold
from lphash_b200 import seqio
from oracle import oracle

CFG1 = os.path.join(GOLDEN_DIR, "config1")
SURVEY_8C = {"self": (4933494, "dcbf27bf3f3ba296"), "salmonella": (4857420, "12503b3bd07cf737"),
             "ecoli1": (4896104, "f6c1cfd350925ad0"), "srr": (459147, "6b8e312e47bdeb71")}


def expected():
    return json.load(open(os.path.join(CFG1, "expected.json")))


def sha(a, dtype="<u8"):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


def test_expected_matches_survey_table():
    exp = expected()
    assert (exp["k"], exp["m"], exp["kmer_bits"], exp["nkmers"]) == (31, 16, 128, 4933494)
    for name, (n, fold) in SURVEY_8C.items():
        assert exp["queries"][name]["n_codes"] == n
        assert exp["queries"][name]["fnv"] == fold
    assert exp["scan"] == {**exp["scan"], "records": 580515, "n_kmers": 4933494, "mm_count": 4942659}


@pytest.mark.parametrize("name", ["self", "salmonella", "ecoli1", "srr"])
def test_oracle_streaming_query_matches_reference(name):
    """ecoli1 (50 runs of N in one 4.9 Mbase record) and the FASTQ (105 reads with N) take the
    reference's non-ACGT streaming quirk: 73 / 4 spurious codes reproduced."""
    exp = expected()["queries"][name]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, exp["file"]))
    assert (len(offsets) - 1, int(offsets[-1])) == (exp["records"], exp["bases"])
    o = oracle.OracleMphf(os.path.join(CFG1, "se.ust.k31_m16_u128.lph"), 128)
    codes, code_off = o.query_batch(bases, offsets)
    assert len(codes) == exp["n_codes"]
    assert [int(x) for x in codes[:8]] == exp["first_codes"]
    assert sha(codes) == exp["sha256_codes"]
    assert sha(np.diff(code_off)) == exp["sha256_counts"]
    if name == "srr":  # the smallest: also through the (slow, pure-Python) fold of SURVEY 8c
        assert f"{fnv_fold(codes):016x}" == exp["fnv"]


def test_oracle_scan_matches_reference():
    exp = expected()["scan"]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, "se.ust.k31.fa.gz"))
    rec, nk, mm = oracle.scan(bases, offsets, 31, 16, mode=0)
    assert (len(rec), nk, mm) == (exp["records"], exp["n_kmers"], exp["mm_count"])
    assert hashlib.sha256(rec.tobytes()).hexdigest() == exp["sha256_records"]


@pytest.mark.parametrize("name", ["ecoli1", "srr"])
def test_oracle_build_side_on_input_with_non_acgt_bytes(name):
    """from_string / classify / get_colliding_kmers over files WITH invalid bytes (the reference flushes the open
    super-k-mer at each one and restarts the window, include/minimizer.hpp:138-151, 283-300)"""
    exp = expected()["scan_dirty"][name]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, expected()["queries"][name]["file"]))
    rec, nk, mm = oracle.scan(bases, offsets, 31, 16, mode=0)
    assert (len(rec), nk, mm) == (exp["records"], exp["n_kmers"], exp["mm_count"])
    assert hashlib.sha256(rec.tobytes()).hexdigest() == exp["sha256_records"]
    trip, ids = oracle.classify(rec)
    assert hashlib.sha256(trip.tobytes()).hexdigest() == exp["sha256_triplets"]
    assert sha(ids) == exp["sha256_ids"]
    km = oracle.colliding_kmers(bases, offsets, 31, 16, ids, kmer_bits=128)
    assert len(km) == exp["colliding_kmers"]
    assert hashlib.sha256(np.ascontiguousarray(km).tobytes()).hexdigest() == exp["sha256_colliding_kmers"]
