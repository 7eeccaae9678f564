"""The device-side loader (image_decode.cu): compact_vector::access, Elias-Fano access, rs_bit_vector::rank,
quartet_wtree::rank_of and the per-bucket part of mphf::query evaluated on the GPU at load time, against the host-side
decode of lph_image.cpp (LPHB_HOST_DECODE=1): the two flat device images must be equal byte for byte - partitioned
and unpartitioned files, both kmer_t flavours, 32- and 64-bit bucket words.  (Every other GPU test loads through the
device-side path, so all code parity rests on it too.)"""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN_DIR, GOLDEN_NAMES, load_golden
from lphash_b200 import api

pytestmark = pytest.mark.gpu
CASES = [(n, False) for n in GOLDEN_NAMES] + [("alt_k31_m20_u64", True), ("alt_k63_m24_u128", True), ("alt_k25_m13_u64", True)]


def load(path, bits, alt):
    return api.Mphf.load_alt(path, bits) if alt else api.Mphf.load(path, bits)


@pytest.mark.parametrize("name,alt", CASES)
@pytest.mark.parametrize("wide", [False, True])
def test_device_decode_equals_host_decode(name, alt, wide, monkeypatch):
    bits = 128 if "u128" in name else 64
    path = os.path.join(GOLDEN_DIR, name + ".lph")
    if wide:
        monkeypatch.setenv("LPHB_FORCE_WIDE_BUCKETS", "1")
    monkeypatch.setenv("LPHB_HOST_DECODE", "1")
    f = load(path, bits, alt)
    want = f.device_image()
    host_info = (f.info.nkmers, f.info.distinct_minimizers, f.info.fallback_keys, f.info.device_bytes)
    f.close()
    monkeypatch.delenv("LPHB_HOST_DECODE")
    f = load(path, bits, alt)
    got = f.device_image()
    assert (f.info.nkmers, f.info.distinct_minimizers, f.info.fallback_keys, f.info.device_bytes) == host_info
    assert len(got) == len(want)
    if got != want:
        a, b = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
        first = int(np.flatnonzero(a != b)[0])
        pytest.fail(f"images differ from byte {first} of {len(got)} ({int((a != b).sum())} bytes)")
    # the one header value that needs a decoded entry (collision_base): the codes of a colliding minimizer depend on it
    g = load_golden(name[4:] if alt else name)
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz")) if alt else g
    codes, _ = f.query_batch(g.q_bases, g.q_offsets)
    assert np.array_equal(codes, z["q_codes"] if alt else g.q_codes)
    f.close()


def test_config1_index_device_decode_equals_host_decode(monkeypatch):
    path = os.path.join(GOLDEN_DIR, "config1", "se.ust.k31_m16_u128.lph")
    monkeypatch.setenv("LPHB_HOST_DECODE", "1")
    f = api.Mphf.load(path, 128)
    want = f.device_image()
    f.close()
    monkeypatch.delenv("LPHB_HOST_DECODE")
    f = api.Mphf.load(path, 128)
    assert f.device_image() == want
    f.close()


@pytest.mark.parametrize("host", [False, True])
def test_inconsistencies_only_decoding_reveals_are_format_errors(host, monkeypatch):
    """structure intact, content wrong: a flipped root bit of the wavelet tree (ones no longer match the max/none
    leaf), an Elias-Fano high bit cleared (fewer set bits than values)"""
    if host:
        monkeypatch.setenv("LPHB_HOST_DECODE", "1")
    image = bytearray(open(os.path.join(GOLDEN_DIR, "k31_m20_u64.lph"), "rb").read())
    sec = api.lph_sections(bytes(image), 64)
    bad = bytearray(image)
    bad[sec[1] + 16] ^= 1  # first word of the root bits (after nbits, nwords)
    with pytest.raises(api.LphashError) as e:
        api.Mphf.from_bytes(bytes(bad), 64)
    assert e.value.code == api.E_FORMAT
    bad = bytearray(image)
    nwords = struct.unpack_from("<Q", image, sec[2] + 8)[0]
    words = np.frombuffer(bytes(image[sec[2] + 16: sec[2] + 16 + 8 * nwords]), dtype="<u8").copy()
    last = int(np.flatnonzero(words)[-1])
    words[last] = 0  # drops at least one set bit of sizes_and_positions' high bits
    bad[sec[2] + 16: sec[2] + 16 + 8 * nwords] = words.tobytes()
    with pytest.raises(api.LphashError) as e:
        api.Mphf.from_bytes(bytes(bad), 64)
    assert e.value.code == api.E_FORMAT
