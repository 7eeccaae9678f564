#!/usr/bin/env python
"""Golden fixtures for the UNPARTITIONED variant (lphash::mphf_alt, build-u / query-u), generated FROM THE
UNMODIFIED REFERENCE (oracle/_ref, built by oracle/build_ref.sh) like tools/make_golden.py:

  tests/golden/alt_NAME.lph   index built by the reference's build-u (mphf_alt::build + essentials::save) from the
                              index set of tests/golden/NAME.npz
  tests/golden/alt_NAME.npz   q_codes / q_code_offsets: the reference's mphf_alt::operator()(contig, len, true) for
                              every contig of NAME's query batch (members, non-members, non-ACGT contigs, ...);
                              q_codes_ns: its non-streaming branch for the contigs of at least k bytes
Run here:  python tools/make_golden_alt.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lphash_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
for name in ("k31_m20_u64", "k63_m24_u128", "k25_m13_u64"):
    z = np.load(os.path.join(OUT, name + ".npz"))
    k, m, bits = int(z["k"]), int(z["m"]), int(z["bits"])
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "index.fa")
        synth.write_fasta(fa, z["index_bases"], z["index_offsets"])
        lph = os.path.join(OUT, "alt_" + name + ".lph")
        csv = ref.build_alt(fa, k, m, lph, bits=bits, tmp_dir=tmp)
    f = ref.RefMphfAlt(lph, bits)
    raw = z["q_bases"].tobytes()
    off = z["q_offsets"]
    recs = [raw[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    codes = [f.query(r) for r in recs]
    q_code_offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum([len(c) for c in codes], out=q_code_offsets[1:])
    ns = [f.query(r, streaming=False) for r in recs if len(r) >= k]
    np.savez_compressed(os.path.join(OUT, "alt_" + name + ".npz"), q_codes=np.concatenate(codes),
                        q_code_offsets=q_code_offsets, q_codes_ns=np.concatenate(ns), csv=csv)
    print(name, f.kmer_count, "k-mers;", len(recs), "query contigs ->", int(q_code_offsets[-1]), "codes |", csv)
    f.close()
