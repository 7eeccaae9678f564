"""The unpartitioned variant (the reference's lphash::mphf_alt, `build-u` / `query-u`) on the GPU: same kernels,
another probe (include/unpartitioned_mphf.hpp:72-192, src/unpartitioned_mphf.cpp:191-206), against golden
vectors generated from the unmodified reference by tools/make_golden_alt.py.  Bit-exact, non-ACGT quirk included."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, load_golden
from lphash_b200 import api

pytestmark = pytest.mark.gpu
NAMES = ["k31_m20_u64", "k63_m24_u128", "k25_m13_u64"]  # tiled E=1, tiled wide windows, generic kernels


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("wide", [False, True])
def test_query_u_matches_reference_golden(name, wide, monkeypatch):
    g = load_golden(name)
    z = np.load(os.path.join(GOLDEN_DIR, "alt_" + name + ".npz"))
    if wide:
        monkeypatch.setenv("LPHB_FORCE_WIDE_BUCKETS", "1")
    f = api.Mphf.load_alt(os.path.join(GOLDEN_DIR, "alt_" + name + ".lph"), g.bits)
    try:
        assert (f.k, f.m) == (g.k, g.m) and f.get_kmer_count() == int(g.n_kmers)
        codes, code_off = f.query_batch(g.q_bases, g.q_offsets)
        assert np.array_equal(code_off, z["q_code_offsets"])
        assert np.array_equal(codes, z["q_codes"])
        # all members: a minimal perfect hash
        codes, _ = f.query_batch(g.index_bases, g.index_offsets)
        n = f.get_kmer_count()
        assert len(codes) == n and np.array_equal(np.sort(codes), np.arange(n, dtype=np.uint64))
        # non-streaming branch (non-ACGT bytes count as 'A')
        contigs = [c for c in g.contigs() if len(c) >= g.k]
        bases = np.frombuffer(b"".join(contigs), dtype=np.uint8)
        offsets = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
        codes, _ = f.query_batch(bases, offsets, streaming=False)
        assert np.array_equal(codes, z["q_codes_ns"])
        # run-length form
        runs, off2, n_codes = f.query_batch_runs(g.q_bases, g.q_offsets)
        assert np.array_equal(api.expand_runs(runs), z["q_codes"])
    finally:
        f.close()


def test_partitioned_loader_rejects_an_unpartitioned_file():
    g = load_golden("k31_m20_u64")
    with pytest.raises(api.LphashError) as e:
        api.Mphf.load(os.path.join(GOLDEN_DIR, "alt_k31_m20_u64.lph"), g.bits)
    assert e.value.code == api.E_FORMAT


@pytest.mark.parametrize("name,bits", [("k31_m20_u64", 64), ("k63_m24_u128", 128), ("k25_m13_u64", 64)])
def test_build_u_part3_and_writer_reproduce_the_reference_file(name, bits):
    """build-u on the GPU around the reference's two PTHash functions: scan + classify of the index contigs,
    lphb_build_inverted_index_alt (positions + sizes), lphb_lph_assemble_alt - byte for byte the file the reference's
    mphf_alt::build saved (src/unpartitioned_mphf.cpp:31-140)"""
    import struct
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    image = open(os.path.join(GOLDEN_DIR, f"alt_{name}.lph"), "rb").read()
    sec = api.lph_sections(image, bits, alt=True)
    k, m, seed, nkmers, distinct, main = struct.unpack_from("<BBQQQQ", image, 0)
    trip, ids, nk, _ = api.scan_classify(z["index_bases"], z["index_offsets"], k, m, seed)
    assert nk == nkmers and len(trip) == distinct
    mo = image[sec[0]:sec[1]]
    info, body = api.build_inverted_index_alt(mo, trip)
    assert info.num_kmers_in_main_index == main
    assert body == image[sec[1]:sec[3]]
    out = api.lph_assemble_alt(k, m, seed, nkmers, distinct, info, mo, body, image[sec[3]:sec[4]])
    assert out == image
    with pytest.raises(api.LphashError) as e:
        api.build_inverted_index_alt(image[sec[3]:sec[4]], trip)  # the fallback function: other keys
    assert e.value.code == api.E_ARG
