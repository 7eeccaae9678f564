"""build-p Parts 1-4 on the GPU + the `.lph` writer, against the files the unmodified reference wrote.

For every golden index: the index contigs go through lphb_scan_classify (Parts 1 + 2), the triplets and the
reference's own serialized minimizer_order (PTHash construction is out of scope: taken from the golden file)
through lphb_build_inverted_index (Part 3), and lphb_lph_assemble puts the image together with the
reference's fallback_kmer_order.  The result must equal the reference's `.lph` byte for byte: wavelet tree
with its rank directories, Elias-Fano prefix sums with the darray1 select index, the header counters.
(ref: src/partitioned_mphf.cpp:92-106, 163-268; file format include/partitioned_mphf.hpp:204-219)"""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from lphash_b200 import api, seqio

pytestmark = pytest.mark.gpu


def parts(image, bits):
    sec = api.lph_sections(image, bits)
    k, m, seed, nkmers, distinct, n_max, rs, ns, npos = struct.unpack_from("<BBQQQQQQQ", image, 0)
    assert sec[0] == 58
    return dict(k=k, m=m, seed=seed, nkmers=nkmers, distinct=distinct, counters=(n_max, rs, ns, npos),
                minimizer_order=image[sec[0]:sec[1]], body=image[sec[1]:sec[3]], fallback=image[sec[3]:sec[4]],
                wtree_bytes=sec[2] - sec[1])


def rebuild(image, bits, bases, offsets, shuffle=None):
    p = parts(image, bits)
    trip, ids, nk, _ = api.scan_classify(bases, offsets, p["k"], p["m"], p["seed"])
    assert nk == p["nkmers"] and len(trip) == p["distinct"]
    if shuffle is not None:
        trip = trip[np.random.default_rng(shuffle).permutation(len(trip))]
    info, body = api.build_inverted_index(p["k"], p["m"], p["minimizer_order"], trip)
    return p, trip, ids, info, body


def test_inverted_index_and_writer_reproduce_the_reference_file(golden):
    image = open(golden.lph, "rb").read()
    p, trip, ids, info, body = rebuild(image, golden.bits, golden.index_bases, golden.index_offsets)
    assert (info.n_maximal, info.right_coll_sizes_start, info.none_sizes_start, info.none_pos_start) == p["counters"]
    assert info.colliding_minimizers == int((trip["size"] == 0).sum())
    assert info.wtree_bytes == p["wtree_bytes"]
    assert body[: info.wtree_bytes] == p["body"][: info.wtree_bytes]  # quartet_wtree
    assert body == p["body"]                                           # + sizes_and_positions
    out = api.lph_assemble(p["k"], p["m"], p["seed"], p["nkmers"], len(trip), info, p["minimizer_order"], body,
                           p["fallback"])
    assert out == image
    # and the GPU-built image answers queries like the reference's
    f = api.Mphf.from_bytes(out, golden.bits)
    codes, _ = f.query_batch(golden.q_bases, golden.q_offsets)
    f.close()
    assert np.array_equal(codes, golden.q_codes)


def test_triplet_order_does_not_matter(golden):
    image = open(golden.lph, "rb").read()
    p, _, _, info, body = rebuild(image, golden.bits, golden.index_bases, golden.index_offsets, shuffle=7)
    assert body == p["body"]


def test_part4_keys_match_the_fallback_function():
    """the colliding k-mers found on the GPU are exactly the key set the reference's fallback_kmer_order was
    built on: the image maps them one to one onto [0, n)"""
    name = "k31_m20_u64"
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    image = open(os.path.join(GOLDEN_DIR, name + ".lph"), "rb").read()
    p, trip, ids, info, _ = rebuild(image, 64, z["index_bases"], z["index_offsets"])
    km = api.colliding_kmers(z["index_bases"], z["index_offsets"], p["k"], p["m"], ids, p["seed"], 64)
    assert np.array_equal(km, z["coll_kmers"])
    n_fallback = struct.unpack_from("<Q", p["fallback"], 8)[0]  # single_phf: seed, num_keys, ...
    assert len(km) == n_fallback


def test_wrong_function_is_refused():
    name = "k31_m20_u64"
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    image = open(os.path.join(GOLDEN_DIR, name + ".lph"), "rb").read()
    p = parts(image, 64)
    trip = z["triplets"]
    with pytest.raises(api.LphashError) as e:  # built on another key set (and another count)
        api.build_inverted_index(p["k"], p["m"], p["fallback"], trip)
    assert e.value.code == api.E_ARG
    other = parts(open(os.path.join(GOLDEN_DIR, "k21_m11_u64.lph"), "rb").read(), 64)
    n = min(len(trip), other["distinct"])
    if other["distinct"] <= len(trip):  # same count, foreign keys: collisions leave cells unset
        with pytest.raises(api.LphashError) as e:
            api.build_inverted_index(p["k"], p["m"], other["minimizer_order"], trip[:n])
        assert e.value.code == api.E_ARG
    with pytest.raises(api.LphashError) as e:
        api.build_inverted_index(p["k"], p["m"], p["minimizer_order"][:-8], trip)
    assert e.value.code == api.E_FORMAT


def test_capacity_error_reports_the_size():
    name = "k21_m11_u64"
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    image = open(os.path.join(GOLDEN_DIR, name + ".lph"), "rb").read()
    p = parts(image, 64)
    with pytest.raises(api.LphashError) as e:
        api.build_inverted_index(p["k"], p["m"], p["minimizer_order"], z["triplets"], capacity=100)
    assert e.value.code == api.E_CAPACITY
    assert api.lib().lphb_inverted_index_bound(len(z["triplets"])) >= len(p["body"])


def test_bundled_index_config1():
    """BASELINE config 1: the reference's own se.ust.k31.fa.gz index (k=31 m=16, 128-bit kmer_t, 611 unitigs,
    4.9 M k-mers, 0.9 % of them behind colliding minimizers) rebuilt around the reference's two PTHash functions"""
    cfg1 = os.path.join(GOLDEN_DIR, "config1")
    image = open(os.path.join(cfg1, "se.ust.k31_m16_u128.lph"), "rb").read()
    bases, offsets = seqio.read_batch(os.path.join(cfg1, "se.ust.k31.fa.gz"))
    p, trip, ids, info, body = rebuild(image, 128, bases, offsets)
    assert body == p["body"]
    out = api.lph_assemble(p["k"], p["m"], p["seed"], p["nkmers"], len(trip), info, p["minimizer_order"], body,
                           p["fallback"])
    assert out == image


@pytest.mark.parametrize("name", ["sparse", "maximal", "mixed"])
def test_synthetic_branches_match_the_restatement(name):
    """low width 0 and sparse darray1 blocks (overflow positions), an empty sizes_and_positions, every type mixed
    over 70,000 minimizers: the device result equals oracle/invindex.py, which tests/test_oracle_golden.py pins on
    these same inputs against the reference's own ef_sequence (tools/make_golden_part3.py)"""
    import hashlib
    from conftest import PART3_FIXTURES, part3_triplets
    from oracle import invindex, oracle
    z = np.load(os.path.join(GOLDEN_DIR, f"part3_{name}.npz"))
    n, k, m = PART3_FIXTURES[name]
    trip = part3_triplets(name)
    phf = z["minimizer_order"].tobytes()
    info, body = api.build_inverted_index(k, m, phf, trip)
    assert [info.n_maximal, info.right_coll_sizes_start, info.none_sizes_start, info.none_pos_start] == \
        [int(v) for v in z["counters"]]
    assert len(body) == int(z["body_bytes"])
    if hashlib.sha256(body).hexdigest() != str(z["body_sha256"]):  # locate the first difference for the report
        _, want = invindex.build_inverted_index(trip, oracle.phf_positions(phf, trip["itself"]), k, m)
        first = next(i for i in range(len(want)) if want[i] != body[i])
        pytest.fail(f"body differs from the restatement at byte {first} (wtree {info.wtree_bytes} bytes)")


# ---- build input with non-ACGT bytes (include/minimizer.hpp:138-151) --------------------------------
def dirty_build_batch(k, m, seed):
    """contigs whose valid runs hit every case of the reference's loop: runs shorter than m / than k, of exactly
    k bases followed by an invalid byte (k-mer counted, never emitted), of exactly k bases at a contig end
    (emitted), long runs, long runs of N, contigs that start / end with N, empty and all-N contigs; a few runs are
    repeated so that some minimizers collide (ids for get_colliding_kmers)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGTacgt", dtype=np.uint8)

    def run(n):
        return rng.choice(acgt[:4] if rng.random() < 0.8 else acgt, int(n))

    N, junk = np.array([78], np.uint8), np.array([45], np.uint8)
    rep = run(3 * k)
    contigs = [
        np.concatenate([run(100), N, run(k), N, run(50), N, N, run(k - 1), junk, run(k)]),
        run(200),
        np.concatenate([N, run(k), junk, run(k + 1), np.repeat(N, 5 * k + 3), run(500), N]),
        np.repeat(N, 10), np.zeros(0, np.uint8), run(k), run(k - 1),
        np.concatenate([run(m - 1), N, run(m), N, rep, N, run(2 * k), N, rep, np.repeat(N, k - 1), rep[: 2 * k]]),
        np.concatenate([run(4000), N, run(3000), junk, run(k), N]),
        run(2500),
    ]
    bases = np.concatenate(contigs).astype(np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
    return bases, offsets


@pytest.mark.parametrize("k,m", [(31, 20), (63, 24), (15, 7), (25, 13)])  # fused scan kernel (W <= 17, wide) / generic passes
def test_build_input_with_non_acgt_bytes_matches_the_reference_loop(k, m):
    from oracle import oracle
    bases, offsets = dirty_build_batch(k, m, seed=k)
    want, wk, wmm = oracle.scan(bases, offsets, k, m, mode=0, mm_count=5)
    got, gk, gmm = api.scan_superkmers(bases, offsets, k, m, mm_count=5)
    assert (gk, gmm) == (wk, wmm)
    assert np.array_equal(got, want)
    # Parts 1 + 2 fused, and Part 4 on the same input
    want0, _, wmm0 = oracle.scan(bases, offsets, k, m, mode=0)
    trip_w, ids_w = oracle.classify(want0)
    trip, ids, nk, mm = api.scan_classify(bases, offsets, k, m)
    assert nk == wk and mm == wmm0 and np.array_equal(trip, trip_w) and np.array_equal(ids, ids_w)
    assert len(ids) > 0
    bits = 128 if k > 31 else 64
    km_w = oracle.colliding_kmers(bases, offsets, k, m, ids_w, kmer_bits=bits)
    km = api.colliding_kmers(bases, offsets, k, m, ids, kmer_bits=bits)
    assert np.array_equal(km, km_w)
    # a batch that does not start at offset 0, clean contigs after a dirty call
    sub = offsets[2:7]
    want2, wk2, _ = oracle.scan(bases[int(sub[0]):int(sub[-1])], sub - sub[0], k, m, mode=0)
    got2, gk2, _ = api.scan_superkmers(bases, sub, k, m)
    assert gk2 == wk2 and np.array_equal(got2, want2)
    clean, coff = bases[int(offsets[1]):int(offsets[2])], np.array([0, offsets[2] - offsets[1]], np.uint64)
    want3, _, _ = oracle.scan(clean, coff, k, m, mode=0)
    got3, _, _ = api.scan_superkmers(clean, coff, k, m)
    assert np.array_equal(got3, want3)


@pytest.mark.parametrize("k,m", [(31, 20), (25, 13), (63, 24)])
def test_invalid_bytes_in_contigs_too_short_for_a_kmer_still_count(k, m):
    """m-mer ordinals advance over valid runs only (include/minimizer.hpp:45-49), also in contigs that hold no k-mer
    and therefore never reach the scan kernels: the ids of every later record depend on them.  (Found by
    tools/stress_parity.py: the generic kernels missed an `R` in a 24-base contig.)"""
    from oracle import oracle
    rng = np.random.default_rng(k + m)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    short = rng.choice(acgt, k - 1)
    short[2] = ord("R")                      # m <= length < k, one invalid byte: 5 fewer m-mers than its length says
    short2 = rng.choice(acgt, m + 1)
    short2[m // 2] = ord("*")                # two runs shorter than m: no m-mer at all
    contigs = [short, rng.choice(acgt, 400), short2, rng.choice(acgt, k), short.copy()]
    bases = np.concatenate(contigs).astype(np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
    want, wk, wmm = oracle.scan(bases, offsets, k, m, mode=0, mm_count=3)
    got, gk, gmm = api.scan_superkmers(bases, offsets, k, m, mm_count=3)
    assert (gk, gmm) == (wk, wmm) and np.array_equal(got, want)
    # a batch without a single k-mer: nothing to scan, the ordinals still move by the valid m-mers only
    only_short = np.concatenate([short, short2, short]).astype(np.uint8)
    off2 = np.array([0, len(short), len(short) + len(short2), len(only_short)], dtype=np.uint64)
    want2, wk2, wmm2 = oracle.scan(only_short, off2, k, m, mode=0, mm_count=7)
    got2, gk2, gmm2 = api.scan_superkmers(only_short, off2, k, m, mm_count=7)
    assert (len(got2), gk2, gmm2) == (len(want2), wk2, wmm2) == (0, 0, wmm2)
