"""The literal drop-in: the reference's OWN command-line driver (src/lphash.cpp + src/query.cpp + src/build.cpp,
compiled unmodified by integration/dropin/build_dropin.sh against a shadow of include/partitioned_mphf.hpp) with
`lphash::mphf` answering on the GPU, run beside the reference's CLI on the reference's bundled data.

query-p prints `query file, mphf file, #k-mers, ns/k-mer streaming, ns/k-mer random` (src/query.cpp:83-86): the
k-mer count (the streaming pass's; the non-streaming pass must agree or the reference's own assert would fire in
a debug build) has to equal the reference CLI's and the committed expectation.  build-p runs Parts 1-4 on the GPU
(scan, sort + classify, inverted index, colliding k-mers; the two PTHash constructions stay the reference's own
CPU calls) and must write the very file and print the very CSV line the reference CLI does; --check then drives the
GPU-backed class through check_collisions / check_streaming_correctness / check_perfection
(include/mphf_utils.hpp:51-100), i.e. both branches of operator() contig by contig."""
import json
import os
import subprocess
import time

import pytest

from conftest import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu
CFG1 = os.path.join(GOLDEN_DIR, "config1")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "lphash128")
GPU_CLI = os.path.join(ROOT, "oracle", "_ref", "lphash_gpu128")
LPH = os.path.join(CFG1, "se.ust.k31_m16_u128.lph")


def need_binaries():
    if not (os.path.exists(REF_CLI) and os.path.exists(GPU_CLI)):
        pytest.skip("drop-in binaries not built (integration/dropin/build_dropin.sh needs the reference tree)")


def cli_env(extra=None):
    """the library's own defaults (the suite pins LPHB_GENERIC_BELOW for its in-process tests, conftest.py)"""
    env = {k: v for k, v in os.environ.items() if k != "LPHB_GENERIC_BELOW"}
    env.update(extra or {})
    return env


def run_query(cli, path):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "query-p", "-i", LPH, "-q", path], capture_output=True, text=True, timeout=900, env=cli_env())
    secs = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
    f = r.stdout.strip().splitlines()[-1].split(",")
    return int(f[2]), float(f[3]), float(f[4]), secs


@pytest.mark.parametrize("name", ["self", "salmonella"])
def test_query_p_counts_equal_the_reference_cli(name):
    need_binaries()
    exp = json.load(open(os.path.join(CFG1, "expected.json")))["queries"][name]
    path = os.path.join(CFG1, exp["file"])
    n_ref, s_ref, r_ref, wall_ref = run_query(REF_CLI, path)
    n_gpu, s_gpu, r_gpu, wall_gpu = run_query(GPU_CLI, path)
    assert n_gpu == n_ref == exp["n_codes"]
    print(f"\n{name}: reference CLI {s_ref:.1f} / {r_ref:.1f} ns per k-mer (streaming / random), wall {wall_ref:.2f} s; "
          f"GPU drop-in {s_gpu:.1f} / {r_gpu:.1f} ns per k-mer, wall {wall_gpu:.2f} s")


def test_query_p_on_a_file_with_non_acgt_bytes():
    """The streaming pass reproduces the reference's quirk (count 459,147 on the FASTQ), the non-streaming pass
    counts every window (L-k+1 per read): the CSV carries the streaming count, as the reference's does."""
    need_binaries()
    exp = json.load(open(os.path.join(CFG1, "expected.json")))["queries"]["srr"]
    path = os.path.join(CFG1, exp["file"])
    n_ref = run_query(REF_CLI, path)[0]
    n_gpu = run_query(GPU_CLI, path)[0]
    assert n_gpu == n_ref == exp["n_codes"]


def run_build(cli, src, k, m, out, tmp, extra=(), env=None):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "build-p", "-i", src, "-k", str(k), "-m", str(m), "-o", out, "-d", tmp, *extra],
                       capture_output=True, text=True, timeout=1800, env=cli_env(env))
    assert r.returncode == 0, r.stderr
    return r.stdout, r.stderr, time.perf_counter() - t0


def test_build_p_on_the_gpu_writes_the_reference_file_and_passes_check(tmp_path):
    need_binaries()
    src = os.path.join(CFG1, "se.ust.k31.fa.gz")
    out_gpu, out_ref = str(tmp_path / "gpu.lph"), str(tmp_path / "ref.lph")
    csv_gpu, err_gpu, wall_gpu = run_build(GPU_CLI, src, 31, 16, out_gpu, str(tmp_path), ["--check"])
    assert "Everything is ok" in err_gpu, err_gpu[-2000:]
    csv_ref, _, wall_ref = run_build(REF_CLI, src, 31, 16, out_ref, str(tmp_path))
    assert csv_gpu == csv_ref  # input file, k, m, collision rate, density bounds, bits per k-mer
    image = open(out_gpu, "rb").read()
    assert image == open(out_ref, "rb").read() == open(LPH, "rb").read()
    print(f"\nbuild-p se.ust.k31 (4.9 M k-mers): reference CLI {wall_ref:.2f} s, GPU drop-in incl. --check {wall_gpu:.2f} s")


def test_build_p_64_bit_flavour_and_cpu_switch(tmp_path):
    """the uint64_t kmer_t binary on a small FASTA; LPHASH_B200_CPU_BUILD=1 forwards the build to the reference"""
    gpu64, ref64 = GPU_CLI.replace("128", "64"), REF_CLI.replace("128", "64")
    if not (os.path.exists(gpu64) and os.path.exists(ref64)):
        pytest.skip("64-bit drop-in binaries not built")
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, "k31_m20_u64.npz"))
    raw, off = z["index_bases"].tobytes(), z["index_offsets"]
    fa = tmp_path / "index.fa"
    with open(fa, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">%d\n" % i + raw[int(off[i]):int(off[i + 1])] + b"\n")
    outs = {}
    for tag, cli, env in [("gpu", gpu64, None), ("ref", ref64, None),
                          ("gpu_second_attempt", gpu64, {"LPHASH_B200_FIRST_CAP": "10"}),  # triplet buffer too small at first
                          ("cpu_switch", gpu64, {"LPHASH_B200_CPU_BUILD": "1"})]:
        out = str(tmp_path / (tag + ".lph"))
        csv, _, _ = run_build(cli, str(fa), 31, 20, out, str(tmp_path), env=env)
        outs[tag] = (csv, open(out, "rb").read())
    assert outs["gpu"] == outs["ref"] == outs["cpu_switch"] == outs["gpu_second_attempt"]
    assert outs["gpu"][1] == open(os.path.join(GOLDEN_DIR, "k31_m20_u64.lph"), "rb").read()


def test_build_u_and_query_u_through_the_gpu_classes(tmp_path):
    """the unpartitioned twin: `build-u` on the GPU writes the reference CLI's file and CSV line and passes --check;
    `query-u` counts the same k-mers"""
    gpu64, ref64 = GPU_CLI.replace("128", "64"), REF_CLI.replace("128", "64")
    if not (os.path.exists(gpu64) and os.path.exists(ref64)):
        pytest.skip("64-bit drop-in binaries not built")
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, "k31_m20_u64.npz"))
    raw, off = z["index_bases"].tobytes(), z["index_offsets"]
    fa = tmp_path / "index.fa"
    with open(fa, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">%d\n" % i + raw[int(off[i]):int(off[i + 1])] + b"\n")
    outs = {}
    for tag, cli, extra in [("gpu", gpu64, ["--check"]), ("ref", ref64, [])]:
        out = str(tmp_path / (tag + ".lph"))
        r = subprocess.run([cli, "build-u", "-i", str(fa), "-k", "31", "-m", "20", "-o", out, "-d", str(tmp_path), *extra],
                           capture_output=True, text=True, timeout=900, env=cli_env())
        assert r.returncode == 0, r.stderr
        if extra:
            assert "Everything is ok" in r.stderr, r.stderr[-2000:]
        outs[tag] = (r.stdout, open(out, "rb").read())
    assert outs["gpu"] == outs["ref"]
    assert outs["gpu"][1] == open(os.path.join(GOLDEN_DIR, "alt_k31_m20_u64.lph"), "rb").read()
    counts = []
    for cli in (gpu64, ref64):
        r = subprocess.run([cli, "query-u", "-i", str(tmp_path / "gpu.lph"), "-q", str(fa)], capture_output=True, text=True,
                           timeout=900, env=cli_env())
        assert r.returncode == 0, r.stderr
        counts.append(int(r.stdout.strip().splitlines()[-1].split(",")[2]))
    assert counts[0] == counts[1] == int(z["n_kmers"])
