"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares,
rejects malformed images before touching CUDA, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT, load_golden
from lphash_b200 import api

HEADER = os.path.join(ROOT, "include", "lphash_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lphb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(api.EXPORTS) == names


def test_version_and_error_string():
    L = api.lib()
    assert b"sm_100a" in L.lphb_version()
    assert isinstance(L.lphb_last_error(), bytes)


def _load_bytes(image: bytes, bits=64):
    h = C.c_void_p()
    buf = (C.c_char * max(len(image), 1)).from_buffer_copy(image or b"\0")
    rc = api.lib().lphb_mphf_load_memory(C.addressof(buf), len(image), bits, 0, C.byref(h))
    return rc, h


def test_missing_file_is_io_error():
    h = C.c_void_p()
    rc = api.lib().lphb_mphf_load_file(b"/nonexistent/x.lph", 64, 0, C.byref(h))
    assert rc == api.E_IO and not h.value


@pytest.mark.parametrize("name", ["k31_m20_u64", "k63_m24_u128"])
def test_malformed_images_are_format_errors(name):
    g = load_golden(name)
    image = open(g.lph, "rb").read()
    for bad in (image[:-1], image[: len(image) // 2], image[:40], image + b"\0", b""):
        rc, h = _load_bytes(bad, g.bits)
        assert rc == api.E_FORMAT, api.lib().lphb_last_error()
        assert not h.value
    # k too large for a 64-bit kmer_t
    if g.k > 31:
        rc, h = _load_bytes(image, 64)
        assert rc == api.E_FORMAT
    rc, h = _load_bytes(image, 32)
    assert rc == api.E_FORMAT


def test_no_gpu_means_loud_failure_not_fallback():
    """On a box without a CUDA device a well-formed image must fail with E_CUDA."""
    n = C.c_int(-1)
    rc = api.lib().lphb_device_count(C.byref(n))
    if rc == api.OK and n.value > 0:
        pytest.skip("a GPU is present")
    g = load_golden("k31_m20_u64")
    with pytest.raises(api.LphashError) as e:
        api.Mphf.load(g.lph, g.bits)
    assert e.value.code == api.E_CUDA
    with pytest.raises(api.LphashError) as e:
        api.scan_superkmers(g.index_bases, g.index_offsets, g.k, g.m)
    assert e.value.code == api.E_CUDA


def test_argument_errors():
    L = api.lib()
    assert L.lphb_mphf_load_file(None, 64, 0, None) == api.E_ARG
    assert L.lphb_mphf_info(None, None) == api.E_ARG
    nrec, nk, mm = C.c_uint64(), C.c_uint64(), C.c_uint64()
    off = np.zeros(1, dtype=np.uint64)
    rc = L.lphb_scan_superkmers(0, 31, 40, 42, None, off.ctypes.data, 0, C.byref(mm), None, 0,
                                C.byref(nrec), C.byref(nk))
    assert rc in (api.E_ARG, api.E_CUDA)


def test_expand_runs_is_host_only_and_exact():
    """lphb_expand_runs decodes run records on the host (no device needed): ascending, descending,
    single codes, wrap-around mod 2^64, many threads, capacity error."""
    import ctypes as C
    rng = np.random.Generator(np.random.PCG64(5))
    n_runs = 50000
    runs = np.zeros(n_runs, dtype=api.RUN_DTYPE)
    runs["first"] = rng.integers(0, 1 << 63, size=n_runs, dtype=np.uint64)
    runs["n"] = rng.integers(-40, 41, size=n_runs)
    runs["n"][runs["n"] == 0] = 1
    runs["first"][0], runs["n"][0] = 2, -5           # 2, 1, 0, 2^64-1, 2^64-2 (mod 2^64 like the reference)
    runs["first"][1], runs["n"][1] = (1 << 64) - 2, 4
    want = []
    for f, n in zip(runs["first"].tolist(), runs["n"].tolist()):
        step = 1 if n > 0 else -1
        want.extend(((f + step * j) & 0xFFFFFFFFFFFFFFFF) for j in range(abs(n)))
    want = np.array(want, dtype=np.uint64)
    for threads in (1, 3, 16):
        got = api.expand_runs(runs, threads=threads)
        assert np.array_equal(got, want), threads
    assert len(api.expand_runs(runs[:0])) == 0
    small = np.empty(3, dtype=np.uint64)
    n = C.c_uint64(0)
    rc = api.lib().lphb_expand_runs(runs.ctypes.data, n_runs, small.ctypes.data, 3, C.byref(n), 1)
    assert rc == api.E_CAPACITY and n.value == len(want)


def test_unpartitioned_images_parse_and_cross_loading_is_rejected():
    """The host-side parser of the mphf_alt format runs before any CUDA call: on a box without a GPU a good
    image gets as far as LPHB_E_CUDA, a partitioned image fed to the unpartitioned loader (or truncated) stops
    at LPHB_E_FORMAT."""
    import ctypes as C
    L = api.lib()
    h = C.c_void_p()
    alt = os.path.join(GOLDEN_DIR, "alt_k31_m20_u64.lph")
    rc = L.lphb_mphf_alt_load_file(alt.encode(), 64, 0, C.byref(h))
    assert rc in (0, api.E_CUDA), L.lphb_last_error()
    if rc == 0:
        L.lphb_mphf_free(h)
    part = os.path.join(GOLDEN_DIR, "k31_m20_u64.lph")
    assert L.lphb_mphf_alt_load_file(part.encode(), 64, 0, C.byref(h)) == api.E_FORMAT
    assert L.lphb_mphf_load_file(alt.encode(), 64, 0, C.byref(h)) == api.E_FORMAT
    img = open(alt, "rb").read()
    for cut in (10, 58, len(img) // 2, len(img) - 1):
        buf = (C.c_char * cut).from_buffer_copy(img[:cut])
        assert L.lphb_mphf_alt_load_memory(C.addressof(buf), cut, 64, 0, C.byref(h)) == api.E_FORMAT


def test_lph_sections_and_assemble_round_trip_on_the_host():
    """host-only entry points: cutting a reference-written file into its parts (lphb_lph_sections) and putting it
    together again (lphb_lph_assemble / _alt) gives the same bytes; sizes that do not add up are refused"""
    import struct
    for name, bits, alt in [("k31_m20_u64", 64, False), ("k63_m24_u128", 128, False), ("alt_k31_m20_u64", 64, True)]:
        image = open(os.path.join(GOLDEN_DIR, name + ".lph"), "rb").read()
        sec = api.lph_sections(image, bits, alt)
        assert sec[0] == (34 if alt else 58) and sec[4] == len(image) and sec == sorted(sec)
        mo, body, fb = image[sec[0]:sec[1]], image[sec[1]:sec[3]], image[sec[3]:sec[4]]
        if alt:
            k, m, seed, nkmers, distinct, main = struct.unpack_from("<BBQQQQ", image, 0)
            idx = api.InvertedIndexAlt(num_kmers_in_main_index=main, positions_bytes=sec[2] - sec[1], sizes_bytes=sec[3] - sec[2])
            out = api.lph_assemble_alt(k, m, seed, nkmers, distinct, idx, mo, body, fb)
        else:
            k, m, seed, nkmers, distinct, n_max, rs, ns, npos = struct.unpack_from("<BBQQQQQQQ", image, 0)
            idx = api.InvertedIndex(n_maximal=n_max, right_coll_sizes_start=rs, none_sizes_start=ns, none_pos_start=npos,
                                    wtree_bytes=sec[2] - sec[1], ef_bytes=sec[3] - sec[2])
            out = api.lph_assemble(k, m, seed, nkmers, distinct, idx, mo, body, fb)
            with pytest.raises(api.LphashError) as e:
                api.lph_assemble(k, m, seed, nkmers, distinct, idx, mo, body[:-8], fb)
            assert e.value.code == api.E_ARG
        assert out == image
    with pytest.raises(api.LphashError) as e:
        api.lph_sections(image[:100], 64, True)
    assert e.value.code == api.E_FORMAT
