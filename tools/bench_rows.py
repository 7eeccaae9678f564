#!/usr/bin/env python
"""Measurements of the SURVEY.md §8 rows that bench.py's headline line does not cover, one JSON line
each (kept under profiles/):

  scan    build-side minimizer / super-k-mer scan (lphb_scan_superkmers) over the config-2 unitigs,
          k=31 m=20, host buffers in and out (the C ABI has no device-resident scan entry), parity
          against the oracle on a prefix, the reference's from_string timed beside it
  k63     BASELINE config 4: k=63 m=24, 128-bit kmer_t, streaming query of the index's own unitigs
  reads   BASELINE config 5 shape on one GPU: 150-base reads, half of them substrings of the indexed
          genome with 1 % substitutions, half random (mixed member / non-member k-mers), against the
          config-2 index; parity against the oracle on a prefix

  classify  sort by minimizer + minimizer::classify (lphb_classify) over the scan's record stream
  part3     build-p Part 3 (lphb_build_inverted_index) + the `.lph` writer on the config-2 index: triplets from
            lphb_scan_classify, the reference's own two PTHash functions, result compared byte for byte with the
            file the reference's build-p wrote

    python tools/bench_rows.py [scan] [classify] [k63] [reads] [--kmers N] [--reads N]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402  (workload cache, log)


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


def build_index(tag, bases, offsets, k, m, bits):
    from lphash_b200 import synth
    from oracle import ref
    os.makedirs(B.CACHE, exist_ok=True)
    lph = os.path.join(B.CACHE, tag + ".lph")
    if not os.path.exists(lph):
        fa = os.path.join(B.CACHE, tag + ".fa")
        synth.write_fasta(fa, bases, offsets)
        t0 = time.time()
        csv = ref.build(fa, k, m, lph + ".tmp", bits=bits, threads=min(B.host_threads(), 32), tmp_dir=B.CACHE)
        os.replace(lph + ".tmp", lph)
        os.remove(fa)
        B.log(f"[rows] reference build-p {tag}: {csv} ({time.time() - t0:.1f}s)")
    return lph


def time_query(f, bases, offsets, k, steps=20, warmup=3):
    """Device-resident streaming query: returns (n_codes, kernel_ms mean, ms per step, codes)."""
    import torch
    dev = torch.device("cuda:0")
    n_contigs = len(offsets) - 1
    n_kmers = int(np.maximum(np.diff(offsets).astype(np.int64) - k + 1, 0).sum())
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    d_codes = torch.empty(n_kmers, dtype=torch.int64, device=dev)
    d_code_off = torch.empty(n_contigs + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(4, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def step():
        f.query_device(d_bases.data_ptr(), d_off.data_ptr(), offsets, d_codes.data_ptr(), n_kmers,
                       d_code_off.data_ptr(), d_status.data_ptr(), stream.cuda_stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    st = d_status.cpu().numpy()
    assert st[0] == n_kmers and st[1] == 0, f"unexpected status {st}"
    f.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    kms = float(f.stats().kernel_ms)
    codes = d_codes.cpu().numpy().view(np.uint64)
    return n_kmers, kms, ms, codes


def prefix_batch(bases, offsets, n_contigs):
    n_contigs = min(n_contigs, len(offsets) - 1)
    off = offsets[: n_contigs + 1]
    return bases[int(off[0]): int(off[-1])], off - off[0]


def row_query(name, workload, lph, bits, k, m, bases, offsets, members_all, oracle_contigs, steps):
    from lphash_b200 import api
    from oracle import oracle
    f = api.Mphf.load(lph, bits, device=0)
    n_kmers, kms, ms, codes = time_query(f, bases, offsets, k, steps=steps)
    checks = []
    if members_all:  # an all-member set: codes are a permutation of 0..n-1
        assert len(codes) == f.get_kmer_count() and int(codes.max()) == len(codes) - 1
        chk = np.zeros(len(codes), dtype=np.uint8)
        chk[codes] = 1
        assert int(chk.sum()) == len(codes), "codes are not a permutation"
        checks.append("codes are a permutation of 0..n-1")
    pb, po = prefix_batch(bases, offsets, oracle_contigs)
    o = oracle.OracleMphf(lph, bits)
    want, _ = o.query_batch(pb, po)
    o.close()
    assert np.array_equal(codes[: len(want)], want), "codes differ from the oracle"
    checks.append(f"first {len(want)} codes bit-exact vs the CPU oracle")
    # the reference's own streaming query on the host cores, bounded sample
    cpu = None
    try:
        from oracle import ref
        threads = B.host_threads()
        r = ref.RefMphf(lph, bits)
        sb, so = prefix_batch(bases, offsets, max(64, (len(offsets) - 1) // 4))
        secs, n, _, _ = r.query_batch(sb, so, threads=threads, want_codes=False)
        secs, n, _, _ = r.query_batch(sb, so, threads=threads, want_codes=False)
        r.close()
        cpu = {"value": n / secs, "unit": "k-mers/s", "cores": threads, "kind": "reference",
               "sample": f"{n} k-mers (first quarter of the contigs), one pass after one warm-up"}
    except Exception as e:
        cpu = {"unavailable": str(e)}
    f.close()
    peak, src = peak_gbs()
    algo = int(offsets[-1] - offsets[0]) + 8 * n_kmers
    ach = algo / (kms * 1e-3) / 1e9
    return {"row": name, "metric": "query-p k-mers/sec", "value": n_kmers / (ms * 1e-3), "unit": "k-mers/s",
            "n_gpus": 1, "ms_per_step": ms, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload, "k": k, "m": m, "kmer_bits": bits, "kmers": n_kmers},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": algo, "peak_source": src},
            "parity": checks, "cpu_baseline": cpu}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rows", nargs="*", default=["scan", "k63", "reads"])
    ap.add_argument("--kmers", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from lphash_b200 import api, synth

    if "scan" in args.rows:
        from oracle import oracle
        k, m = 31, 20
        bases, offsets = synth.unitigs(args.kmers, k, m)
        import ctypes as C
        import torch
        # pinned host buffers (what a caller streaming batches would use), straight through the C ABI
        n_contigs = len(offsets) - 1
        cap = int(np.maximum(np.diff(offsets).astype(np.int64) - k + 1, 0).sum()) + 1
        if cap > 1_000_000_000:  # full-scale sets: records are ~2/(w+1) per k-mer, not one (18 B each, pinned)
            cap = cap // 4 + 1_000_000
        h_bases = torch.from_numpy(bases).pin_memory()
        h_rec = torch.empty(cap * 18, dtype=torch.uint8).pin_memory()
        L = api.lib()

        def call():
            mmc, nrec, nkm = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
            rc = L.lphb_scan_superkmers(0, k, m, 42, h_bases.data_ptr(), offsets.ctypes.data, n_contigs, C.byref(mmc),
                                        h_rec.data_ptr(), cap, C.byref(nrec), C.byref(nkm))
            assert rc == 0, L.lphb_last_error()
            return nrec.value, nkm.value, mmc.value

        call()  # warm-up: context, module load, workspace allocation
        times = []
        for _ in range(5 if args.kmers <= 200_000_000 else 2):
            t0 = time.perf_counter()
            nrec, nk, mm = call()
            times.append(time.perf_counter() - t0)
        secs = float(np.mean(times))
        rec = h_rec.numpy()[: nrec * 18].view(api.RECORD_DTYPE)
        pb, po = prefix_batch(bases, offsets, 256)
        wrec, wnk, wmm = oracle.scan(pb, po, k, m, mode=0)
        assert np.array_equal(rec[: len(wrec)], wrec), "scan records differ from the oracle"
        assert int(rec["size"].astype(np.int64).sum()) == nk, "record sizes do not add up to the k-mer count"
        cpu = None
        try:
            from oracle import ref
            sb, so = prefix_batch(bases, offsets, (len(offsets) - 1) // 8)
            t1 = time.perf_counter()
            out = ref.scan(sb, so, k, m, bits=64)
            t1 = time.perf_counter() - t1
            n_s = int(np.maximum(np.diff(so).astype(np.int64) - k + 1, 0).sum())
            cpu = {"value": n_s / t1, "unit": "k-mers/s", "cores": 1, "kind": "reference",
                   "sample": f"{n_s} k-mers (first eighth of the contigs), minimizer::from_string, 1 thread, "
                             "incl. the harness's copy of the records"}
        except Exception as e:
            cpu = {"unavailable": str(e)}
        peak, src = peak_gbs()
        algo = int(offsets[-1]) + 18 * len(rec)
        # device-resident: bases and offsets already in HBM, records left in HBM (lphb_scan_superkmers_device);
        # the kernel time is the CUDA-event time of the record-producing kernel on its stream
        d_bases = torch.from_numpy(bases).cuda()
        d_off = torch.from_numpy(offsets.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        kms, wall = [], []
        drec = None
        for it in range(8):
            t0 = time.perf_counter()
            drec, dn, dk, dmm, ms = api.scan_superkmers_device(d_bases.data_ptr(), d_off.data_ptr(), offsets, k, m,
                                                              fetch=(it == 0))
            wall.append(time.perf_counter() - t0)
            if it == 0:
                assert np.array_equal(drec, rec), "device-resident scan differs from the host-buffer scan"
            elif it >= 3:
                kms.append(ms)
        kern_ms = float(np.mean(kms))
        full = None
        try:  # every record against the unmodified reference's from_string (one thread, ~20 ns per k-mer)
            from oracle import ref
            if nk <= 300_000_000:
                t1 = time.perf_counter()
                wr, wk2, wm2 = ref.scan(bases, offsets, k, m, bits=64)
                t1 = time.perf_counter() - t1
                assert (wk2, wm2) == (nk, mm) and np.array_equal(wr, rec), "records differ from the reference's"
                full = f"all {len(wr)} records equal the unmodified reference's minimizer::from_string stream ({t1:.1f} s on one thread)"
        except (ImportError, OSError, RuntimeError) as e:
            full = f"reference compare unavailable: {e}"
        print(json.dumps({"row": "scan_device", "metric": "build-p scan k-mers/sec", "value": nk / (kern_ms * 1e-3),
                          "unit": "k-mers/s", "n_gpus": 1, "ms_per_step": kern_ms, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": f"synthetic unitigs ({args.kmers} k-mers), build-side minimizer/super-k-mer scan, "
                                                 "bases resident in HBM, records left in HBM (lphb_scan_superkmers_device)",
                                     "k": k, "m": m, "kmers": int(nk), "records": int(len(rec))},
                          "call_ms_incl_host_sync": float(np.mean(wall[3:])) * 1e3,
                          "roofline": {"bound": "hbm", "achieved": algo / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / (kern_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": algo,
                                       "kernel": "k_query_tiled<31,20,scan> (CUDA events on its stream, mean of 5 launches)",
                                       "kernel_ms": kern_ms, "peak_source": src},
                          "parity": [full, f"first {len(wrec)} records bit-exact vs the CPU oracle"]}), flush=True)
        del d_bases, d_off
        print(json.dumps({"row": "scan", "metric": "build-p scan k-mers/sec", "value": nk / secs, "unit": "k-mers/s",
                          "n_gpus": 1, "ms_per_step": secs * 1e3, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": f"synthetic unitigs of BASELINE config {2 if args.kmers <= 200_000_000 else 3} ({args.kmers} k-mers), build-side minimizer/super-k-mer scan "
                                                 f"through lphb_scan_superkmers (pinned host buffers in and out, mean of {len(times)} calls)",
                                     "k": k, "m": m, "kmers": int(nk), "records": int(len(rec)), "mm_count": int(mm)},
                          "e2e": {"value": nk / secs, "unit": "k-mers/s", "h2d_bytes_per_step": int(offsets[-1]) + 8 * len(offsets),
                                  "d2h_bytes_per_step": 18 * len(rec)},
                          "roofline": {"bound": "hbm", "achieved": algo / secs / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / secs / 1e9 / peak, "algorithmic_bytes_per_launch": algo,
                                       "peak_source": src, "note": "whole call incl. PCIe copies (no device-resident entry point)"},
                          "parity": [f"first {len(wrec)} records bit-exact vs the CPU oracle", "sum of record sizes == k-mer count"],
                          "cpu_baseline": cpu}), flush=True)

    if "classify" in args.rows:
        import ctypes as C
        import torch
        from oracle import oracle
        k, m = 31, 20
        bases, offsets = synth.unitigs(args.kmers, k, m)
        rec, nk, mm = api.scan_superkmers(bases, offsets, k, m)  # the GPU scan's record stream, scan order
        n = len(rec)
        h_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).pin_memory()
        h_trip = torch.empty(n * 10, dtype=torch.uint8).pin_memory()
        h_ids = torch.empty(max(n // 8, 1024), dtype=torch.int64).pin_memory()
        L = api.lib()

        def call():
            nt, ni = C.c_uint64(0), C.c_uint64(0)
            rc = L.lphb_classify(0, h_rec.data_ptr(), n, h_trip.data_ptr(), n, C.byref(nt), h_ids.data_ptr(),
                                 len(h_ids), C.byref(ni))
            assert rc == 0, L.lphb_last_error()
            return nt.value, ni.value

        call()
        times = []
        for _ in range(5):
            t0 = time.perf_counter()
            nt, ni = call()
            times.append(time.perf_counter() - t0)
        secs = float(np.mean(times))
        t0 = time.perf_counter()
        want_t, want_i = oracle.classify(rec)
        cpu_secs = time.perf_counter() - t0
        got_t = h_trip.numpy()[: nt * 10].view(api.TRIPLET_DTYPE)
        got_i = h_ids.numpy()[:ni].view(np.uint64)
        assert np.array_equal(got_t, want_t) and np.array_equal(got_i, want_i), "classify differs from the oracle"
        # Parts 1 + 2 fused (records never leave the device): bases in, triplets + ids out
        n_contigs = len(offsets) - 1
        h_bases = torch.from_numpy(bases).pin_memory()

        def fused():
            mmc, nt2, ni2, nk2 = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
            rc = L.lphb_scan_classify(0, k, m, 42, h_bases.data_ptr(), offsets.ctypes.data, n_contigs, C.byref(mmc),
                                      h_trip.data_ptr(), n, C.byref(nt2), h_ids.data_ptr(), len(h_ids), C.byref(ni2),
                                      C.byref(nk2))
            assert rc == 0, L.lphb_last_error()
            return nt2.value, ni2.value, nk2.value

        fused()
        ftimes = []
        for _ in range(5):
            t0 = time.perf_counter()
            nt2, ni2, nk2 = fused()
            ftimes.append(time.perf_counter() - t0)
        fsecs = float(np.mean(ftimes))
        assert (nt2, ni2, nk2) == (nt, ni, nk)
        assert np.array_equal(h_trip.numpy()[: nt * 10].view(api.TRIPLET_DTYPE), want_t)
        assert np.array_equal(h_ids.numpy()[:ni].view(np.uint64), want_i)
        peak, src = peak_gbs()
        print(json.dumps({"row": "scan_classify", "metric": "build-p scan+sort+classify k-mers/sec", "value": nk / fsecs,
                          "unit": "k-mers/s", "n_gpus": 1, "ms_per_step": fsecs * 1e3, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": "config-2 unitigs (k=31 m=20): lphb_scan_classify = from_string stream + sort by "
                                                 "minimizer + classify, records kept on the device (pinned host buffers, mean of 5 calls)",
                                     "kmers": int(nk), "records": int(n), "triplets": int(nt), "colliding_ids": int(ni)},
                          "e2e": {"value": nk / fsecs, "unit": "k-mers/s", "h2d_bytes_per_step": int(offsets[-1]) + 8 * len(offsets),
                                  "d2h_bytes_per_step": 10 * nt + 8 * ni},
                          "parity": ["all triplets and colliding ids bit-exact vs the CPU oracle"]}), flush=True)
        algo = 18 * n + 10 * nt + 8 * ni
        print(json.dumps({"row": "classify", "metric": "build-p sort+classify records/sec", "value": n / secs,
                          "unit": "records/s", "n_gpus": 1, "ms_per_step": secs * 1e3, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": "records of the config-2 unitig scan (k=31 m=20) in scan order, sort by "
                                                 "minimizer + classify through lphb_classify (pinned host buffers, mean of 5 calls)",
                                     "records": int(n), "triplets": int(nt), "colliding_ids": int(ni)},
                          "e2e": {"value": n / secs, "unit": "records/s", "h2d_bytes_per_step": 18 * n,
                                  "d2h_bytes_per_step": 10 * nt + 8 * ni},
                          "roofline": {"bound": "hbm", "achieved": algo / secs / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / secs / 1e9 / peak, "algorithmic_bytes_per_launch": int(algo),
                                       "peak_source": src, "note": "whole call incl. PCIe copies and per-call allocations"},
                          "parity": ["all triplets and colliding ids bit-exact vs the CPU oracle"],
                          "cpu_baseline": {"value": n / cpu_secs, "unit": "records/s", "cores": 1, "kind": "port",
                                           "sample": f"all {n} records, oracle classify (std::stable_sort + one pass), 1 thread"}}),
              flush=True)

    if "part3" in args.rows:
        import struct
        k, m = 31, 20
        bases, offsets, lph = B.make_workload(args.kmers)
        image = open(lph, "rb").read()
        sec = api.lph_sections(image, 64)
        seed, nkmers, distinct = struct.unpack_from("<QQQ", image, 2)
        t0 = time.perf_counter()
        trip, ids, nk, _ = api.scan_classify(bases, offsets, k, m, seed)
        parts12 = time.perf_counter() - t0
        assert nk == nkmers and len(trip) == distinct
        mo = image[sec[0]:sec[1]]
        info, body = api.build_inverted_index(k, m, mo, trip)  # warm-up (context, allocations)
        times, dev = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            info, body = api.build_inverted_index(k, m, mo, trip)
            times.append(time.perf_counter() - t0)
            dev.append(info.device_ms)
        secs, dev_ms = float(np.mean(times)), float(np.mean(dev))
        assert body == image[sec[1]:sec[3]], "inverted index differs from the reference's file"
        t0 = time.perf_counter()
        out = api.lph_assemble(k, m, seed, nkmers, distinct, info, mo, body, image[sec[3]:sec[4]])
        asm = time.perf_counter() - t0
        assert out == image, "assembled .lph differs from the reference's file"
        algo = 10 * distinct + len(body)  # triplets in, serialized index out
        peak, src = peak_gbs()
        print(json.dumps({"row": "part3", "metric": "build-p Part 3 distinct minimizers/sec", "value": distinct / secs,
                          "unit": "minimizers/s", "n_gpus": 1, "ms_per_step": secs * 1e3, "dtype": "u64", "data": "synthetic",
                          "config": {"workload": "config-2 index (k=31 m=20): re-key by minimizer_order + build_inverted_index "
                                                 "(wavelet tree, rank directories, Elias-Fano + darray1) through "
                                                 "lphb_build_inverted_index, pageable host buffers, mean of 3 calls",
                                     "distinct_minimizers": int(distinct), "kmers": int(nkmers), "body_bytes": len(body),
                                     "minimizer_order_bytes": len(mo), "colliding_minimizers": int(info.colliding_minimizers)},
                          "device_ms": dev_ms, "parts_1_2_ms": parts12 * 1e3, "assemble_ms": asm * 1e3,
                          "roofline": {"bound": "hbm", "achieved": algo / (dev_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / (dev_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(algo),
                                       "peak_source": src, "note": "all Part-3 kernels together (CUDA events), copies excluded"},
                          "parity": ["wavelet tree + sizes_and_positions byte-identical to the reference's .lph",
                                     "assembled .lph byte-identical to the reference's file"]}), flush=True)

    if "k63" in args.rows:
        k, m, bits = 63, 24, 128
        bases, offsets = synth.unitigs(args.kmers, k, m, seed=0x5EED0004)
        lph = build_index(f"cfg4_n{args.kmers}_k{k}_m{m}_u{bits}", bases, offsets, k, m, bits)
        print(json.dumps(row_query("k63", "BASELINE config 4: synthetic unitigs (62-base overlaps), all members, "
                                   "128-bit kmer_t, streaming query-p", lph, bits, k, m, bases, offsets,
                                   True, 32, args.steps)), flush=True)

    if "reads" in args.rows:
        k, m, bits = B.K, B.M, B.BITS
        ubases, uoffsets, lph = B.make_workload(args.kmers)
        # the genome the unitigs were cut from = the unitigs with their (k-1)-base overlaps removed; any long
        # stretch of it serves as a source of member substrings
        genome = ubases[: int(uoffsets[-1])]
        rb, ro = synth.reads(args.reads, genome)
        print(json.dumps(row_query("reads", f"BASELINE config 5 shape on one GPU: {args.reads} reads of 150 bases, "
                                   "half substrings of the indexed unitigs with 1% substitutions, half random "
                                   "(mixed member / non-member k-mers), config-2 index", lph, bits, k, m, rb, ro,
                                   False, 20000, args.steps)), flush=True)


if __name__ == "__main__":
    main()
