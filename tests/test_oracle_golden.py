"""The CPU oracle (oracle/lphash_oracle.cpp) against the committed golden vectors that
tools/make_golden.py generated from the UNMODIFIED reference.  Runs without a GPU and without
/root/reference."""
import os

import numpy as np
import pytest

from oracle import oracle


def test_known_answers():
    # SURVEY.md §8 a3/a7 KATs (extracted from the reference)
    assert oracle.murmur64(0, 42) == 7729525510718446440
    assert oracle.murmur64(1, 42) == 7191219851201888087
    assert oracle.murmur64(0xFFFFFFFF, 42) == 2352765155560249676
    assert oracle.murmur64(7, 1) == 12982362733409533052
    assert oracle.fastmod(12345678901234567, 614718) == 103303 == 12345678901234567 % 614718


def test_header_fields(golden):
    o = oracle.OracleMphf(golden.lph, golden.bits)
    assert (o.k, o.m) == (golden.k, golden.m)
    assert o.info["mm_seed"] == 42
    assert o.info["nkmers"] == int(golden.n_kmers)
    assert o.info["file_bytes"] == __import__("os").path.getsize(golden.lph)
    assert o.info["mo_num_keys"] == o.info["distinct_minimizers"] == len(golden.triplets)
    assert o.info["fb_num_keys"] == len(golden.coll_kmers)


def test_streaming_query_matches_reference(golden):
    """operator()(contig, len, true) — every contig incl. the ones with non-ACGT bytes (Q1)."""
    o = oracle.OracleMphf(golden.lph, golden.bits)
    codes, code_off = o.query_batch(golden.q_bases, golden.q_offsets)
    assert np.array_equal(code_off, golden.q_code_offsets)
    assert np.array_equal(codes, golden.q_codes)


def test_stateless_equals_streaming_on_clean_contigs(golden):
    """S1: on ACGT-only contigs the stateless definition (what the GPU runs) == streaming."""
    o = oracle.OracleMphf(golden.lph, golden.bits)
    off = golden.q_code_offsets
    n_clean = 0
    for i, (c, clean) in enumerate(zip(golden.contigs(), golden.is_clean())):
        want = golden.q_codes[int(off[i]):int(off[i + 1])]
        got = o.query_stateless(c)
        if clean:
            n_clean += 1
            assert len(want) == max(0, len(c) - golden.k + 1)
            assert np.array_equal(got, want), i
        else:
            # valid k-mers only: a subsequence of the reference's output (spurious entries removed)
            assert len(got) <= len(want) or len(want) == 0
    assert n_clean > 100


def test_index_set_is_a_minimal_perfect_hash(golden):
    o = oracle.OracleMphf(golden.lph, golden.bits)
    codes, _ = o.query_batch(golden.index_bases, golden.index_offsets)
    n = int(golden.n_kmers)
    assert len(codes) == n
    assert np.array_equal(np.sort(codes), np.arange(n, dtype=np.uint64))


def test_scan_matches_from_string(golden):
    for mode in (0, 1):  # sequential restatement, stateless definition
        rec, nk, mm = oracle.scan(golden.index_bases, golden.index_offsets, golden.k, golden.m, mode=mode)
        assert nk == int(golden.n_kmers) and mm == int(golden.mm_count)
        assert np.array_equal(rec, golden.rec), mode


def test_scan_sequential_on_dirty_and_short_contigs_runs(golden):
    # the sequential restatement accepts anything; sizes must add up on clean contigs
    rec, nk, _ = oracle.scan(golden.index_bases, golden.index_offsets, golden.k, golden.m, mode=0)
    assert int(rec["size"].astype(np.int64).sum()) == nk


def test_classify_and_colliding_kmers(golden):
    trip, ids = oracle.classify(golden.rec)
    assert np.array_equal(trip, golden.triplets)
    assert np.array_equal(ids, golden.coll_ids)
    km = oracle.colliding_kmers(golden.index_bases, golden.index_offsets, golden.k, golden.m,
                                golden.coll_ids, kmer_bits=golden.bits)
    assert km.shape == golden.coll_kmers.shape
    assert np.array_equal(km, golden.coll_kmers)


# ---- build-p Part 3 (oracle/invindex.py) ------------------------------------------------------------
def _sections(image, bits):
    from lphash_b200 import api  # host-only entry point (parser of the serialized image)
    return api.lph_sections(image, bits)


def test_part3_restatement_reproduces_the_reference_files(golden):
    """re-key + build_inverted_index restated on the CPU = the wavelet tree and sizes_and_positions bytes of the
    `.lph` the unmodified reference wrote (src/partitioned_mphf.cpp:92-106, 163-268)"""
    import struct
    from oracle import invindex
    image = open(golden.lph, "rb").read()
    sec = _sections(image, golden.bits)
    trip = golden.triplets
    orders = oracle.phf_positions(image[sec[0]:sec[1]], trip["itself"])
    counters, body = invindex.build_inverted_index(trip, orders, golden.k, golden.m)
    assert counters == struct.unpack_from("<QQQQ", image, 26)
    assert body == image[sec[1]:sec[3]]


@pytest.mark.parametrize("name", ["sparse", "maximal", "mixed"])
def test_part3_restatement_on_synthetic_branches(name):
    """low width 0 + sparse darray1 blocks with overflow positions, an empty sequence, every type mixed: the
    Elias-Fano image equals what the reference's own ef_sequence::encode saved (tools/make_golden_part3.py)"""
    import hashlib
    from conftest import GOLDEN_DIR, PART3_FIXTURES, part3_triplets
    from oracle import invindex
    z = np.load(os.path.join(GOLDEN_DIR, f"part3_{name}.npz"))
    n, k, m = PART3_FIXTURES[name]
    trip = part3_triplets(name)
    orders = oracle.phf_positions(z["minimizer_order"].tobytes(), trip["itself"])
    assert np.array_equal(np.sort(orders), np.arange(n, dtype=np.uint64))
    counters, body = invindex.build_inverted_index(trip, orders, k, m)
    assert list(counters) == [int(v) for v in z["counters"]]
    assert len(body) == int(z["body_bytes"]) and hashlib.sha256(body).hexdigest() == str(z["body_sha256"])
    ef = invindex.ef_sequence_image(np.cumsum(invindex.value_lists(trip, orders, k, m)).tolist(),
                                    int(invindex.value_lists(trip, orders, k, m).sum()))
    assert hashlib.sha256(ef).hexdigest() == str(z["ef_sha256"])


@pytest.mark.parametrize("name,bits", [("k31_m20_u64", 64), ("k63_m24_u128", 128), ("k25_m13_u64", 64)])
def test_part3_restatement_unpartitioned(name, bits):
    """build-u: positions + sizes of the `.lph` the reference's mphf_alt::build saved (src/unpartitioned_mphf.cpp:78-96,
    152-169)"""
    import struct
    from conftest import GOLDEN_DIR, load_golden
    from oracle import invindex
    image = open(os.path.join(GOLDEN_DIR, f"alt_{name}.lph"), "rb").read()
    from lphash_b200 import api
    sec = api.lph_sections(image, bits, alt=True)
    trip = load_golden(name).triplets
    orders = oracle.phf_positions(image[sec[0]:sec[1]], trip["itself"])
    main, body = invindex.build_inverted_index_alt(trip, orders)
    assert main == struct.unpack_from("<Q", image, 26)[0]
    assert body == image[sec[1]:sec[3]]
