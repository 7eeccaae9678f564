#!/usr/bin/env python
"""Randomised parity stress (one-off, not part of the test suite): for every golden index, many random batches -
substrings of the index sequence with substitutions, random sequence, lower case, N and junk bytes, lengths from 0
to a few tiles, random batch sizes and base offsets - through the CUDA path (streaming query, run-length form,
build scan, classify, colliding k-mers) against the CPU oracle.  usage: stress_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN_NAMES, load_golden  # noqa: E402
from lphash_b200 import api  # noqa: E402
from oracle import oracle  # noqa: E402


def random_batch(rng, g):
    idx = g.index_bases
    big = rng.random() < 0.08  # now and then a batch of a few hundred tiles (several CTAs per SM, many chunks)
    n = int(rng.integers(200, 1500)) if big else int(rng.integers(1, 60))
    pieces = []
    for _ in range(n):
        kind = rng.integers(0, 10)
        L = int(rng.choice([0, 1, g.m - 1, g.m, g.k - 1, g.k, g.k + 1, int(rng.integers(0, 200)), int(rng.integers(200, 3000)),
                            int(rng.integers(900, 1100))]))
        if big and rng.random() < 0.02:
            L = int(rng.integers(5000, 40000))
        if kind < 5 and L < len(idx):
            s0 = int(rng.integers(0, len(idx) - L))
            p = idx[s0:s0 + L].copy()
            if kind < 2 and L:
                where = rng.integers(0, L, size=max(1, L // 80))
                p[where] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), len(where))
        elif kind == 5 and L:
            # low complexity: homopolymers and short-period repeats (every window full of equal m-mers: the leftmost
            # one wins, partitioned_mphf.hpp:124,152,159; hash-prefix ties everywhere)
            unit = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(1, 7)))
            p = np.tile(unit, L // len(unit) + 1)[:L].copy()
            if L > 50 and rng.random() < 0.5:  # a few point changes inside the repeat
                p[rng.integers(0, L, size=3)] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 3)
        else:
            p = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)
        if kind == 7 and L:
            p[rng.integers(0, L, size=int(rng.integers(1, 4)))] = rng.choice(np.frombuffer(b"N-nRY*", dtype=np.uint8))
        if kind == 8 and L > 40:
            a = int(rng.integers(0, L - 30))
            p[a:a + int(rng.integers(1, 30))] = ord("N")
        if kind == 9:
            p = np.frombuffer(p.tobytes().lower(), dtype=np.uint8)
        pieces.append(p.astype(np.uint8))
    bases = np.concatenate(pieces) if pieces else np.zeros(0, np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in pieces])]).astype(np.uint64)
    return bases, offsets


def main_alt(seconds, seed):
    """the unpartitioned variant (mphf_alt, query-u) and, beside it, the partitioned one, against the UNMODIFIED
    REFERENCE itself (oracle/_ref travels to the GPU box): streaming and non-streaming branch, record by record"""
    from oracle import ref
    rng = np.random.default_rng(seed)
    cases = []
    for name in ("k31_m20_u64", "k63_m24_u128", "k25_m13_u64"):
        g = load_golden(name)
        alt = os.path.join(os.path.dirname(g.lph), "alt_" + name + ".lph")
        cases.append((g, api.Mphf.load_alt(alt, g.bits), ref.RefMphfAlt(alt, g.bits)))
        cases.append((g, api.Mphf.load(g.lph, g.bits), ref.RefMphf(g.lph, g.bits)))
    t_end = time.time() + seconds
    rounds = 0
    while time.time() < t_end:
        for g, f, r in cases:
            bases, offsets = random_batch(rng, g)
            raw = bases.tobytes()
            recs = [raw[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]
            for streaming in (True, False):
                want = [r.query(c, streaming) for c in recs if streaming or len(c) >= g.k]
                want = np.concatenate(want) if want else np.zeros(0, np.uint64)
                got, _ = f.query_batch(bases, offsets, streaming=streaming)
                if not np.array_equal(got, want):
                    path = os.path.join(ROOT, "gpurun_out", f"stress_fail_alt_{g.name}_{rounds}.npz")
                    np.savez_compressed(path, bases=bases, offsets=offsets)
                    print(f"MISMATCH vs the reference ({type(r).__name__}, streaming={streaming}) {g.name}: saved to {path}", flush=True)
                    return 1
        rounds += 1
    print(f"stress vs the reference ok: {rounds} rounds x {len(cases)} functions (3 mphf_alt, 3 mphf), seed {seed}", flush=True)
    return 0


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if len(sys.argv) > 3 and sys.argv[3] == "alt":
        return main_alt(seconds, seed)
    rng = np.random.default_rng(seed)
    handles = {n: api.Mphf.load(load_golden(n).lph, load_golden(n).bits) for n in GOLDEN_NAMES}
    os.environ["LPHB_FORCE_WIDE_BUCKETS"] = "1"  # second set of handles: 64-bit bucket words
    wide = {n: api.Mphf.load(load_golden(n).lph, load_golden(n).bits) for n in GOLDEN_NAMES}
    del os.environ["LPHB_FORCE_WIDE_BUCKETS"]
    oracles = {n: oracle.OracleMphf(load_golden(n).lph, load_golden(n).bits) for n in GOLDEN_NAMES}
    valid = np.zeros(256, dtype=bool)
    for ch in b"ACGTUacgtu":
        valid[ch] = True
    t_end = time.time() + seconds
    rounds = 0
    while time.time() < t_end:
        for name in GOLDEN_NAMES:
            g = load_golden(name)
            bases, offsets = random_batch(rng, g)
            os.environ["LPHB_GENERIC_BELOW"] = str(int(rng.choice([0, 1 << 18])))
            if rng.random() < 0.3:
                os.environ["LPHB_NO_SMALL_PATH"] = "1"
            else:
                os.environ.pop("LPHB_NO_SMALL_PATH", None)
            f, o = (wide if rng.random() < 0.3 else handles)[name], oracles[name]
            os.environ["LPHB_CHUNK_BASES"] = str(int(rng.choice([1, 300, 2000, 1 << 23])))  # chunk pipeline on / off
            want, want_off = o.query_batch(bases, offsets)
            got, got_off = f.query_batch(bases, offsets)
            fails = []
            if not (np.array_equal(got_off, want_off) and np.array_equal(got, want)):
                fails.append("query")
            # a sub-batch whose offsets do not start at zero
            if len(offsets) > 4:
                a, b2 = sorted(int(x) for x in rng.choice(len(offsets), size=2, replace=False))
                if b2 > a:
                    sub = offsets[a:b2 + 1]
                    got_s, off_s = f.query_batch(bases, sub)
                    if not (np.array_equal(got_s, want[int(want_off[a]):int(want_off[b2])]) and
                            np.array_equal(off_s, want_off[a:b2 + 1] - want_off[a])):
                        fails.append("sub-batch")
            # non-streaming branch: invalid bytes count as 'A', every window of k bytes answered statelessly
            clean = bases.copy()
            clean[~valid[clean]] = ord("A")
            ns_want = [o.query_stateless(clean[int(offsets[i]):int(offsets[i + 1])].tobytes())
                       for i in range(len(offsets) - 1) if int(offsets[i + 1] - offsets[i]) >= g.k]
            ns_want = np.concatenate(ns_want) if ns_want else np.zeros(0, np.uint64)
            ns_got, _ = f.query_batch(bases, offsets, streaming=False)
            if not np.array_equal(ns_got, ns_want):
                fails.append("non-streaming")
            runs, _, n = f.query_batch_runs(bases, offsets)
            if not np.array_equal(api.expand_runs(runs), want):
                fails.append("runs")
            rec_w, nk_w, mm_w = oracle.scan(bases, offsets, g.k, g.m, mode=0)
            rec, nk, mm = api.scan_superkmers(bases, offsets, g.k, g.m)
            if not ((nk, mm) == (nk_w, mm_w) and np.array_equal(rec, rec_w)):
                fails.append(f"scan nk {nk}/{nk_w} mm {mm}/{mm_w} records {len(rec)}/{len(rec_w)}")
            trip_w, ids_w = oracle.classify(rec_w)
            trip, ids, _, _ = api.scan_classify(bases, offsets, g.k, g.m)
            if not (np.array_equal(trip, trip_w) and np.array_equal(ids, ids_w)):
                fails.append("classify")
            km_w = oracle.colliding_kmers(bases, offsets, g.k, g.m, ids_w, kmer_bits=g.bits)
            km = api.colliding_kmers(bases, offsets, g.k, g.m, ids_w, kmer_bits=g.bits)
            if not np.array_equal(km, km_w):
                fails.append(f"colliding k-mers {len(km)}/{len(km_w)}")
            ok = not fails
            if not ok:
                print("failed:", fails, "env", os.environ.get("LPHB_GENERIC_BELOW"), os.environ.get("LPHB_NO_SMALL_PATH"), flush=True)
                path = os.path.join(ROOT, "gpurun_out", f"stress_fail_{name}_{rounds}.npz")
                os.makedirs(os.path.dirname(path), exist_ok=True)
                np.savez_compressed(path, bases=bases, offsets=offsets)
                print(f"MISMATCH {name} round {rounds}: batch saved to {path}", flush=True)
                return 1
        rounds += 1
    print(f"stress ok: {rounds} rounds x {len(GOLDEN_NAMES)} indexes, seed {seed}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
