#!/usr/bin/env python
"""End to end FROM A FILE: the config-2 unitigs written as FASTA (plain and gzip), queried by
examples/lphb_query.cpp (streaming ingest: one thread inflates + splits records while the GPU works on the previous
chunk) and by the reference's own CLI (`lphash query-p`, gz + kseq + hf per record inside its timed loop,
src/query.cpp:48-56).  One JSON line per run; the folds of all codes must agree between the output forms."""
import gzip
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from lphash_b200 import synth  # noqa: E402


def write_bgzf(src, dst, block=65280, level=1):
    import struct
    import zlib

    def one(chunk):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(payload) + 25) +
                payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))

    with open(src, "rb") as fi, open(dst, "wb") as fo:
        while True:
            chunk = fi.read(block)
            if not chunk:
                break
            fo.write(one(chunk))
        fo.write(one(b""))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    bases, offsets, lph = B.make_workload(n)
    tmp = os.path.join(B.CACHE, "ingest")
    os.makedirs(tmp, exist_ok=True)
    fa = os.path.join(tmp, "q.fa")
    synth.write_fasta(fa, bases, offsets)
    t0 = time.time()
    subprocess.check_call(f"gzip -1 -k -f {fa}", shell=True)
    B.log(f"[ingest] {os.path.getsize(fa)} B of FASTA, {os.path.getsize(fa + '.gz')} B gzip -1 ({time.time() - t0:.0f}s)")
    bgz = fa + ".bgz"  # blocked gzip as bgzip writes it (BGZF): inflated block-parallel by the ingest
    t0 = time.time()
    write_bgzf(fa, bgz)
    B.log(f"[ingest] {os.path.getsize(bgz)} B BGZF ({time.time() - t0:.0f}s)")
    exe = os.path.join(tmp, "lphb_query")
    libdir = os.path.join(ROOT, "lphash_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DLPHASH_B200_WITH_ZLIB", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "lphb_query.cpp"), "-o", exe, "-L", libdir, "-llphash_b200",
                           f"-Wl,-rpath,{libdir}", "-lz", "-pthread"])
    folds = set()
    mb = os.environ.get("INGEST_CHUNK_MB", "4")  # chunk of text per batch (tools: INGEST_CHUNK_SWEEP=1 sweeps it)
    runs_list = [(fa, ["0", mb]), (fa, ["0", mb, "runs"]), (fa + ".gz", ["0", mb]), (fa + ".gz", ["0", mb, "runs"]),
                 (bgz, ["0", mb]), (bgz, ["0", mb, "runs"]),
                 (fa, ["0", mb, "nofold"]), (fa, ["0", mb, "runs", "nofold"]), (fa + ".gz", ["0", mb, "nofold"]),
                 (bgz, ["0", mb, "nofold"]), (bgz, ["0", mb, "runs", "nofold"])]
    if os.environ.get("INGEST_CHUNK_SWEEP"):  # chunk-size sweep: plain text, codes and runs untouched
        runs_list = [(fa, ["0", str(mb)] + extra + ["nofold"]) for mb in (2, 4, 8, 16, 32, 64) for extra in ([], ["runs"])]
    for path, form in runs_list:
        subprocess.check_output([exe, lph, "64", path] + form)  # warm-up: page cache, CUDA context
        t0 = time.perf_counter()
        out = subprocess.check_output([exe, lph, "64", path] + form, text=True).strip().split(",")
        wall = time.perf_counter() - t0
        if "nofold" not in form:
            folds.add(out[5])
        print(json.dumps({"row": "ingest", "impl": "lphash_b200 (examples/lphb_query.cpp, streaming ingest)",
                          "file": os.path.basename(path), "output": "runs" if "runs" in form else "codes",
                          "chunk_mb": int(form[1]) if len(form) > 1 else 64,
                          "host_fold_of_every_code": "nofold" not in form, "kmers": int(out[2]),
                          "ns_per_kmer_end_to_end": float(out[3]), "ns_per_kmer_gpu_calls": float(out[4]),
                          "bases_per_s_end_to_end": float(out[6]), "process_wall_s": wall, "fold": out[5]}), flush=True)
    assert len(folds) == 1, folds
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "lphash64")
    if os.path.exists(ref_cli):
        for path in (fa, fa + ".gz"):
            t0 = time.perf_counter()
            out = subprocess.check_output([ref_cli, "query-p", "-i", lph, "-q", path], text=True).strip().split(",")
            wall = time.perf_counter() - t0
            print(json.dumps({"row": "ingest", "impl": "reference CLI (lphash query-p, unmodified)", "file": os.path.basename(path),
                              "kmers": int(out[2]), "ns_per_kmer_streaming_pass": float(out[3]),
                              "ns_per_kmer_random_pass": float(out[4]), "process_wall_s": wall,
                              "bases_per_s_streaming_pass": (len(bases) / (float(out[3]) * 1e-9 * int(out[2])))}), flush=True)
    for p in (fa, fa + ".gz"):
        os.remove(p)


if __name__ == "__main__":
    main()
