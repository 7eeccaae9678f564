"""The literal drop-in: the reference's OWN command-line driver (src/lphash.cpp + src/query.cpp + src/build.cpp,
compiled unmodified by integration/dropin/build_dropin.sh against a shadow of include/partitioned_mphf.hpp) with
`lphash::mphf` answering on the GPU, run beside the reference's CLI on the reference's bundled data.

query-p prints `query file, mphf file, #k-mers, ns/k-mer streaming, ns/k-mer random` (src/query.cpp:83-86): the
k-mer count (the streaming pass's; the non-streaming pass must agree or the reference's own assert would fire in
a debug build) has to equal the reference CLI's and the committed expectation.  build-p --check drives the
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


def run_query(cli, path):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "query-p", "-i", LPH, "-q", path], capture_output=True, text=True, timeout=900)
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


def test_build_p_check_passes_through_the_gpu_class(tmp_path):
    need_binaries()
    out = str(tmp_path / "se.lph")
    r = subprocess.run([GPU_CLI, "build-p", "-i", os.path.join(CFG1, "se.ust.k31.fa.gz"), "-k", "31", "-m", "16",
                        "-o", out, "-d", str(tmp_path), "--check"], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stderr
    assert "Everything is ok" in r.stderr, r.stderr[-2000:]
    assert open(out, "rb").read() == open(LPH, "rb").read()  # build-p itself is the reference's: same file
