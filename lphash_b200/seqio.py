"""FASTA / FASTQ (optionally gzip) reader producing the batch layout the C ABI consumes:
one uint8 array of concatenated bases + a uint64 offsets array (n_contigs + 1).

Mirrors what kseq.h hands to `mphf::operator()` in the reference driver
(/root/reference/src/query.cpp:51-52): `seq->seq.s` is the record's sequence with line breaks
removed (multi-line FASTA is joined), nothing else is altered (case, N, ... are preserved).
"""
from __future__ import annotations

import gzip

import numpy as np


def _open(path: str):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def read_records(path: str) -> list[bytes]:
    """Sequences of a FASTA/FASTQ file in file order."""
    out: list[bytes] = []
    with _open(path) as f:
        data = f.read()
    lines = data.split(b"\n")
    i, n = 0, len(lines)
    while i < n:
        ln = lines[i]
        if ln.startswith(b">"):
            i += 1
            parts = []
            while i < n and not lines[i].startswith(b">"):
                parts.append(lines[i].rstrip(b"\r"))
                i += 1
            out.append(b"".join(parts))
        elif ln.startswith(b"@"):
            # FASTQ (4-line records, as produced by every modern tool)
            seq = lines[i + 1].rstrip(b"\r") if i + 1 < n else b""
            out.append(seq)
            i += 4
        else:
            i += 1
    return out


def concat(records: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    lens = np.fromiter((len(r) for r in records), dtype=np.uint64, count=len(records))
    offsets = np.zeros(len(records) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    bases = np.frombuffer(b"".join(records), dtype=np.uint8).copy()
    return bases, offsets


def read_batch(path: str) -> tuple[np.ndarray, np.ndarray]:
    return concat(read_records(path))
