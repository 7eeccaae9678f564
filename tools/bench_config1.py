#!/usr/bin/env python
"""BASELINE config 1 on the GPU, timed: the reference's bundled query files (tests/golden/config1/) through
lphb_query_stream with host buffers, beside the clean self-query; every run checked against the unmodified
reference's outputs (expected.json).  ecoli1.fasta (one 4.9 Mbase record, 50 runs of N) and the FASTQ take the
non-ACGT path (quirk_kernels.cu).  One JSON line per file."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lphash_b200 import api, seqio  # noqa: E402

CFG1 = os.path.join(ROOT, "tests", "golden", "config1")
exp = json.load(open(os.path.join(CFG1, "expected.json")))
f = api.Mphf.load(os.path.join(CFG1, "se.ust.k31_m16_u128.lph"), 128)
for name in ("self", "salmonella", "ecoli1", "srr"):
    e = exp["queries"][name]
    bases, offsets = seqio.read_batch(os.path.join(CFG1, e["file"]))
    f.query_batch(bases, offsets)  # warm-up (workspace allocation)
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        codes, code_off = f.query_batch(bases, offsets)
        times.append(time.perf_counter() - t0)
    st = f.stats()
    ok = hashlib.sha256(np.ascontiguousarray(codes, dtype="<u8").tobytes()).hexdigest() == e["sha256_codes"]
    print(json.dumps({"row": "config1_" + name, "file": e["file"], "records": e["records"], "bases": e["bases"],
                      "codes": int(len(codes)), "dirty_contigs": int(st.dirty_contigs), "kernel_launches": int(st.kernel_launches),
                      "ms_per_call_host_buffers": 1e3 * float(np.median(times)),
                      "kmers_per_s_host_buffers": len(codes) / float(np.median(times)),
                      "equal_to_reference": bool(ok)}), flush=True)
    assert ok, name
f.close()
