"""Sharding of the path over ranks (lphash_b200/shard.py): plan / take / mm_count
carry / concatenation are checked on CPU with a world-size-2 gloo group, the per-shard work being
done by the oracle (the checker; there is no CPU product path), and on the GPU with the CUDA
path doing the per-shard work."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from lphash_b200 import shard
from oracle import oracle


def test_plan_covers_everything_and_balances():
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 5000, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64) + np.uint64(17)  # offsets[0] != 0
    for world in (1, 2, 3, 4, 8):
        p = shard.plan(off, world)
        assert len(p) == world and p[0][0] == 0 and p[-1][1] == 1000
        assert all(p[i][1] == p[i + 1][0] for i in range(world - 1))
        sizes = [int(off[b] - off[a]) for a, b in p]
        assert max(sizes) - min(sizes) <= 2 * int(lens.max())


def test_plan_edge_cases():
    assert shard.plan(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4
    p = shard.plan(np.array([0, 10, 20], np.uint64), 8)  # fewer contigs than ranks
    assert p[0][0] == 0 and p[-1][1] == 2 and sum(b - a for a, b in p) == 2
    with pytest.raises(ValueError):
        shard.plan(np.array([0, 1], np.uint64), 0)


def test_sharded_equals_whole_oracle():
    g = load_golden("k31_m20_u64")
    o = oracle.OracleMphf(g.lph, g.bits)
    want, want_off = o.query_batch(g.q_bases, g.q_offsets)
    wrec, wnk, wmm = oracle.scan(g.index_bases, g.index_offsets, g.k, g.m)
    for world in (2, 3, 5):
        p = shard.plan(g.q_offsets, world)
        parts = [o.query_batch(*shard.take(g.q_bases, g.q_offsets, r)) for r in p]
        codes, off = shard.concat_codes(parts)
        assert np.array_equal(codes, want) and np.array_equal(off, want_off)
        p = shard.plan(g.index_offsets, world)
        mm0 = shard.mm_count_starts(g.index_offsets, p, g.m)
        recs, nk, mm = [], 0, 0
        for r, start in zip(p, mm0):
            rec, n, mm = oracle.scan(*shard.take(g.index_bases, g.index_offsets, r), g.k, g.m, mm_count=start)
            recs.append(rec)
            nk += n
        assert nk == wnk and mm == wmm and np.array_equal(shard.concat_records(recs), wrec)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        o = oracle.OracleMphf(g.lph, g.bits)
        p = shard.plan(g.q_offsets, world)
        codes, off = o.query_batch(*shard.take(g.q_bases, g.q_offsets, p[rank]))
        all_codes = shard.gather_to_rank0(codes, dist, rank, world)
        all_off = shard.gather_to_rank0(off, dist, rank, world)
        ps = shard.plan(g.index_offsets, world)
        mm0 = shard.mm_count_starts(g.index_offsets, ps, g.m)
        rec, nk, mm = oracle.scan(*shard.take(g.index_bases, g.index_offsets, ps[rank]), g.k, g.m,
                                  mm_count=mm0[rank])
        all_rec = shard.gather_to_rank0(rec, dist, rank, world)
        dist.barrier()
        if rank == 0:
            got, got_off = shard.concat_codes(list(zip(all_codes, all_off)))
            want, want_off = o.query_batch(g.q_bases, g.q_offsets)
            wrec, _, _ = oracle.scan(g.index_bases, g.index_offsets, g.k, g.m)
            ok = (np.array_equal(got, want) and np.array_equal(got_off, want_off)
                  and np.array_equal(shard.concat_records(all_rec), wrec)
                  and np.array_equal(got, g.q_codes))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, "k31_m20_u64", q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["k31_m20_u64", "k63_m24_u128"])
def test_sharded_equals_whole_gpu(name):
    """Per-shard work on the CUDA path (every shard on the one visible GPU, as separate calls)."""
    from lphash_b200 import api
    g = load_golden(name)
    f = api.Mphf.load(g.lph, g.bits)
    for world in (2, 4):
        p = shard.plan(g.q_offsets, world)
        parts = [f.query_batch(*shard.take(g.q_bases, g.q_offsets, r)) for r in p]
        codes, off = shard.concat_codes(parts)
        assert np.array_equal(codes, g.q_codes) and np.array_equal(off, g.q_code_offsets)
        ps = shard.plan(g.index_offsets, world)
        mm0 = shard.mm_count_starts(g.index_offsets, ps, g.m)
        recs = [api.scan_superkmers(*shard.take(g.index_bases, g.index_offsets, r), g.k, g.m, mm_count=s)[0]
                for r, s in zip(ps, mm0)]
        wrec, _, _ = oracle.scan(g.index_bases, g.index_offsets, g.k, g.m)
        assert np.array_equal(shard.concat_records(recs), wrec)
    f.close()
