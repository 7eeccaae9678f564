"""include/lphash_b200.hpp — the C++ host mirror of lphash::mphf / namespace minimizer — compiled
with g++ against the C-ABI library and run on the golden fixtures (tests/cpp/shim_check.cpp makes
one call per contig, like the reference's driver and builder)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "cpp", "shim_check.cpp")
LIBDIR = os.path.join(ROOT, "lphash_b200")


def build_shim(tmp, bits):
    exe = os.path.join(tmp, f"shim_check{bits}")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", f"-DLPHASH_B200_KMER_BITS={bits}",
                           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-llphash_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def write_batch(path, g):
    def batch(bases, offsets):
        b = bases.tobytes()
        pad = (-len(b)) % 8
        return (np.uint64(len(offsets) - 1).tobytes() + offsets.astype("<u8").tobytes() + b + b"\0" * pad)

    ids = g.coll_ids.astype("<u8")
    with open(path, "wb") as f:
        f.write(batch(g.q_bases, g.q_offsets))
        f.write(np.array([g.k, g.m, len(ids)], dtype="<u8").tobytes() + ids.tobytes())
        f.write(batch(g.index_bases, g.index_offsets))


@pytest.mark.parametrize("bits", [64, 128])
def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path, bits):
    exe = build_shim(str(tmp_path), bits)
    from lphash_b200 import api
    try:
        have_gpu = api.device_count() > 0
    except api.LphashError:
        have_gpu = False
    if have_gpu:
        pytest.skip("a GPU is present")
    g = load_golden("k31_m20_u64" if bits == 64 else "k31_m16_u128")
    write_batch(str(tmp_path / "batch.bin"), g)
    r = subprocess.run([exe, g.lph, str(tmp_path / "batch.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "CUDA" in r.stderr  # std::runtime_error, no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["k31_m20_u64", "k15_m7_u64", "k31_m16_u128", "k63_m24_u128"])
def test_shim_matches_reference_golden(tmp_path, name):
    g = load_golden(name)
    exe = build_shim(str(tmp_path), g.bits)
    write_batch(str(tmp_path / "batch.bin"), g)
    r = subprocess.run([exe, g.lph, str(tmp_path / "batch.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    pos = 0

    def u64():
        nonlocal pos
        v = int(np.frombuffer(raw, dtype="<u8", count=1, offset=pos)[0])
        pos += 8
        return v

    n = u64()
    codes = np.frombuffer(raw, dtype="<u8", count=n, offset=pos)
    pos += 8 * n
    assert np.array_equal(codes, g.q_codes)
    n = u64()
    rec = np.frombuffer(raw, dtype=g.rec.dtype, count=n, offset=pos)
    pos += 18 * n
    assert np.array_equal(rec, g.rec)
    assert u64() == int(g.n_kmers)
    assert u64() == int(g.mm_count)
    n = u64()
    words = g.bits // 64
    coll = np.frombuffer(raw, dtype="<u8", count=n * words, offset=pos).reshape(n, words)
    assert np.array_equal(coll, g.coll_kmers)
    pos += 8 * n * words
    n = u64()
    trip = np.frombuffer(raw, dtype=g.triplets.dtype, count=n, offset=pos)
    assert np.array_equal(trip, g.triplets)
