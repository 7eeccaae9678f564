"""One process, two GPUs, two host threads: every entry point keeps its state per handle or per device (resident-CTA
cache, scan / classify sessions with one mutex per device, Part-3 workspace per call), so the two threads run the
whole path concurrently and independently.  Skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
import threading

import numpy as np
import pytest

from conftest import Golden
from lphash_b200 import api

pytestmark = pytest.mark.gpu


def whole_path(device, name, rounds, errors):
    try:
        g = Golden(name)  # its own NpzFile: the shared cache is not for concurrent readers
        image = open(g.lph, "rb").read()
        sec = api.lph_sections(image, g.bits)
        f = api.Mphf.load(g.lph, g.bits, device=device)
        for _ in range(rounds):
            codes, _ = f.query_batch(g.q_bases, g.q_offsets)
            assert np.array_equal(codes, g.q_codes)
            runs, _, n = f.query_batch_runs(g.q_bases, g.q_offsets)
            assert np.array_equal(api.expand_runs(runs), g.q_codes)
            rec, nk, mm = api.scan_superkmers(g.index_bases, g.index_offsets, g.k, g.m, device=device)
            assert np.array_equal(rec, g.rec)
            trip, ids, nk2, _ = api.scan_classify(g.index_bases, g.index_offsets, g.k, g.m, device=device)
            assert nk2 == nk and np.array_equal(trip, g.triplets) and np.array_equal(ids, g.coll_ids)
            km = api.colliding_kmers(g.index_bases, g.index_offsets, g.k, g.m, ids, kmer_bits=g.bits, device=device)
            assert np.array_equal(km, g.coll_kmers)
            _, body = api.build_inverted_index(g.k, g.m, image[sec[0]:sec[1]], trip, device=device)
            assert body == image[sec[1]:sec[3]]
        f.close()
    except Exception as e:  # noqa: BLE001 - reported by the test below
        errors.append((device, name, repr(e)))


def test_two_devices_two_threads_whole_path():
    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    errors = []
    threads = [threading.Thread(target=whole_path, args=(0, "k31_m20_u64", 6, errors)),
               threading.Thread(target=whole_path, args=(1, "k63_m24_u128", 6, errors)),
               threading.Thread(target=whole_path, args=(1, "k31_m20_u64", 6, errors)),
               threading.Thread(target=whole_path, args=(0, "k25_m13_u64", 6, errors))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
