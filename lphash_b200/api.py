"""Host-side mirror of the reference's `lphash::mphf` query interface on top of the C ABI
(include/lphash_b200.h, built into lphash_b200/liblphash_b200.so).

Reference interface mirrored (/root/reference/include/partitioned_mphf.hpp:13-39):
    hf = mphf(); essentials::load(hf, file)        ->  hf = Mphf.load(file, kmer_bits)
    hf(contig, len, streaming=true)                 ->  hf(contig)            (np.uint64 array)
    hf.get_kmer_count(), hf.get_minimizer_L0()      ->  same names
Batch entry points (`query_batch`, `query_device`) are what a driver that wants throughput calls:
one C-ABI call for many contigs.

There is no CPU implementation behind this module: if the CUDA library is missing or no GPU is
visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LPHASH_B200_LIB selects a tuning variant of the same library (csrc/Makefile: `make variant`)
LIB_PATH = os.environ.get("LPHASH_B200_LIB") or os.path.join(HERE, "liblphash_b200.so")

RECORD_DTYPE = np.dtype([("itself", "<u8"), ("id", "<u8"), ("p1", "u1"), ("size", "u1")])  # mm_record_t

OK, E_ARG, E_IO, E_FORMAT, E_CUDA, E_CAPACITY, E_NOMEM = 0, -1, -2, -3, -4, -5, -6


class LphashError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lphash_b200 error {code}: {msg}")
        self.code = code


class Info(C.Structure):
    _fields_ = [("k", C.c_uint32), ("m", C.c_uint32), ("kmer_bits", C.c_uint32), ("device", C.c_int32),
                ("mm_seed", C.c_uint64), ("nkmers", C.c_uint64), ("distinct_minimizers", C.c_uint64),
                ("n_maximal", C.c_uint64), ("right_coll_sizes_start", C.c_uint64),
                ("none_sizes_start", C.c_uint64), ("none_pos_start", C.c_uint64),
                ("fallback_keys", C.c_uint64), ("file_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("load_host_ms", C.c_double), ("load_h2d_ms", C.c_double)]


class InvertedIndex(C.Structure):
    _fields_ = [("n_maximal", C.c_uint64), ("right_coll_sizes_start", C.c_uint64), ("none_sizes_start", C.c_uint64),
                ("none_pos_start", C.c_uint64), ("colliding_minimizers", C.c_uint64), ("universe", C.c_uint64),
                ("wtree_bytes", C.c_uint64), ("ef_bytes", C.c_uint64), ("device_ms", C.c_double)]


class InvertedIndexAlt(C.Structure):
    _fields_ = [("num_kmers_in_main_index", C.c_uint64), ("positions_bytes", C.c_uint64), ("sizes_bytes", C.c_uint64),
                ("device_ms", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("dirty_contigs", C.c_uint64), ("kernel_ms", C.c_double)]


# every symbol include/lphash_b200.h declares (tests check the library exports all of them)
EXPORTS = ["lphb_last_error", "lphb_version", "lphb_device_count", "lphb_mphf_load_file",
           "lphb_mphf_load_memory", "lphb_mphf_free", "lphb_mphf_info", "lphb_query_stream",
           "lphb_query_stream_device", "lphb_scan_superkmers", "lphb_colliding_kmers",
           "lphb_host_alloc", "lphb_host_free", "lphb_mphf_stats", "lphb_scan_release", "lphb_classify", "lphb_scan_classify",
           "lphb_query_stream_runs", "lphb_expand_runs", "lphb_mphf_dirty_flags",
           "lphb_scan_superkmers_device", "lphb_copy_to_host", "lphb_query_nonstreaming",
           "lphb_mphf_alt_load_file", "lphb_mphf_alt_load_memory", "lphb_inverted_index_bound",
           "lphb_build_inverted_index", "lphb_lph_assemble", "lphb_lph_sections", "lphb_build_inverted_index_alt",
           "lphb_lph_assemble_alt", "lphb_mphf_device_image"]

_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA extension.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
                          " (or make -C lphash_b200/csrc)")
    L = C.CDLL(LIB_PATH)
    u64, p, i32 = C.c_uint64, C.c_void_p, C.c_int
    L.lphb_last_error.restype = C.c_char_p
    L.lphb_version.restype = C.c_char_p
    L.lphb_device_count.argtypes = [C.POINTER(C.c_int)]
    L.lphb_mphf_load_file.argtypes = [C.c_char_p, i32, i32, C.POINTER(p)]
    L.lphb_mphf_load_memory.argtypes = [p, u64, i32, i32, C.POINTER(p)]
    L.lphb_mphf_alt_load_file.argtypes = [C.c_char_p, i32, i32, C.POINTER(p)]
    L.lphb_mphf_alt_load_memory.argtypes = [p, u64, i32, i32, C.POINTER(p)]
    L.lphb_mphf_free.argtypes = [p]
    L.lphb_mphf_info.argtypes = [p, C.POINTER(Info)]
    L.lphb_mphf_stats.argtypes = [p, C.POINTER(Stats)]
    L.lphb_query_stream.argtypes = [p, p, p, u64, p, u64, p, C.POINTER(u64)]
    L.lphb_query_nonstreaming.argtypes = [p, p, p, u64, p, u64, p, C.POINTER(u64)]
    L.lphb_query_stream_device.argtypes = [p, p, p, p, u64, p, u64, p, p, p]
    L.lphb_query_stream_runs.argtypes = [p, p, p, u64, p, u64, C.POINTER(u64), p, C.POINTER(u64)]
    L.lphb_expand_runs.argtypes = [p, u64, p, u64, C.POINTER(u64), i32]
    L.lphb_mphf_dirty_flags.argtypes = [p, C.POINTER(p), C.POINTER(u64)]
    L.lphb_scan_superkmers.argtypes = [i32, C.c_uint32, C.c_uint32, u64, p, p, u64, C.POINTER(u64), p,
                                       u64, C.POINTER(u64), C.POINTER(u64)]
    L.lphb_colliding_kmers.argtypes = [i32, C.c_uint32, C.c_uint32, u64, p, p, u64, C.POINTER(u64), p,
                                       u64, i32, p, u64, C.POINTER(u64)]
    L.lphb_scan_superkmers_device.argtypes = [i32, C.c_uint32, C.c_uint32, u64, p, p, p, u64, C.POINTER(u64), C.POINTER(p),
                                              C.POINTER(u64), C.POINTER(u64), C.POINTER(C.c_double)]
    L.lphb_classify.argtypes = [i32, p, u64, p, u64, C.POINTER(u64), p, u64, C.POINTER(u64)]
    L.lphb_scan_release.argtypes = [i32]
    L.lphb_scan_classify.argtypes = [i32, C.c_uint32, C.c_uint32, u64, p, p, u64, C.POINTER(u64), p, u64,
                                     C.POINTER(u64), p, u64, C.POINTER(u64), C.POINTER(u64)]
    L.lphb_copy_to_host.argtypes = [i32, p, p, u64]
    L.lphb_host_alloc.argtypes = [C.POINTER(p), u64]
    L.lphb_mphf_device_image.argtypes = [p, C.POINTER(p), C.POINTER(u64)]
    L.lphb_inverted_index_bound.argtypes = [u64]
    L.lphb_inverted_index_bound.restype = u64
    L.lphb_build_inverted_index.argtypes = [i32, C.c_uint32, C.c_uint32, p, u64, p, u64, p, u64, C.POINTER(u64),
                                            C.POINTER(InvertedIndex)]
    L.lphb_lph_assemble.argtypes = [C.c_uint32, C.c_uint32, u64, u64, u64, C.POINTER(InvertedIndex), p, u64, p, u64, p,
                                    u64, p, u64, C.POINTER(u64)]
    L.lphb_lph_sections.argtypes = [p, u64, i32, i32, C.POINTER(u64)]
    L.lphb_build_inverted_index_alt.argtypes = [i32, p, u64, p, u64, p, u64, C.POINTER(u64), C.POINTER(InvertedIndexAlt)]
    L.lphb_lph_assemble_alt.argtypes = [C.c_uint32, C.c_uint32, u64, u64, u64, C.POINTER(InvertedIndexAlt), p, u64, p, u64,
                                        p, u64, p, u64, C.POINTER(u64)]
    L.lphb_host_free.argtypes = [p]
    for name in EXPORTS:
        if getattr(L, name).restype is C.c_int:
            pass
    _lib = L
    return L


def _check(rc: int):
    if rc != OK:
        raise LphashError(rc, lib().lphb_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().lphb_device_count(C.byref(n)))
    return n.value


def _as_batch(bases, offsets):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if offsets.ndim != 1 or len(offsets) < 1:
        raise ValueError("offsets must have n_contigs + 1 entries")
    return bases, offsets


class Mphf:
    """A partitioned LP-MPHF resident on one GPU (the reference's `lphash::mphf`, query side)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        info = Info()
        _check(lib().lphb_mphf_info(self._h, C.byref(info)))
        self.info = info
        self.k, self.m, self.kmer_bits, self.device = info.k, info.m, info.kmer_bits, info.device

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def load(cls, path: str, kmer_bits: int = 64, device: int = 0) -> "Mphf":
        """essentials::load(hf, path) (/root/reference/src/query.cpp:37)."""
        h = C.c_void_p()
        _check(lib().lphb_mphf_load_file(os.fsencode(path), kmer_bits, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def load_alt(cls, path: str, kmer_bits: int = 64, device: int = 0) -> "Mphf":
        """An unpartitioned LP-MPHF (the reference's `lphash::mphf_alt`, build-u / query-u)."""
        h = C.c_void_p()
        _check(lib().lphb_mphf_alt_load_file(os.fsencode(path), kmer_bits, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_bytes(cls, image: bytes, kmer_bits: int = 64, device: int = 0) -> "Mphf":
        h = C.c_void_p()
        buf = (C.c_char * len(image)).from_buffer_copy(image)
        _check(lib().lphb_mphf_load_memory(C.addressof(buf), len(image), kmer_bits, device, C.byref(h)))
        return cls(h.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().lphb_mphf_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference-named accessors -----------------------------------------------------------
    def get_kmer_count(self) -> int:
        return self.info.nkmers

    def get_minimizer_L0(self) -> int:
        return self.info.distinct_minimizers

    def device_image(self) -> bytes:
        """the flat device image behind the handle, copied to the host (diagnostic)"""
        ptr, n = C.c_void_p(), C.c_uint64(0)
        _check(lib().lphb_mphf_device_image(self._h, C.byref(ptr), C.byref(n)))
        out = np.empty(n.value, dtype=np.uint8)
        _check(lib().lphb_copy_to_host(self.info.device, out.ctypes.data, ptr, n.value))
        return out.tobytes()

    def dirty_flags(self) -> np.ndarray:
        """per contig of the last query call: nonzero where the contig holds a byte outside ACGT/acgt/U/u"""
        ptr, n = C.c_void_p(), C.c_uint64(0)
        _check(lib().lphb_mphf_dirty_flags(self._h, C.byref(ptr), C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            _check(lib().lphb_copy_to_host(self.info.device, out.ctypes.data, ptr, n.value))
        return out

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().lphb_mphf_stats(self._h, C.byref(s)))
        return s

    # -- queries -----------------------------------------------------------------------------
    def __call__(self, contig, streaming: bool = True) -> np.ndarray:
        """hf(contig, len, streaming): hash codes of the contig's k-mers.  The two modes of the
        reference agree on ACGT-only input; `streaming=False` is the reference's
        non-streaming branch, where a non-ACGT byte counts as 'A' (lphb_query_nonstreaming)."""
        if isinstance(contig, str):
            contig = contig.encode()
        bases = np.frombuffer(bytes(contig), dtype=np.uint8)
        offsets = np.array([0, len(bases)], dtype=np.uint64)
        codes, _ = self.query_batch(bases, offsets, streaming=streaming)
        return codes

    def query_batch(self, bases, offsets, out: np.ndarray | None = None, streaming: bool = True):
        """One C-ABI call for a batch: returns (codes, code_offsets) with HOST arrays in and out.
        streaming=False: the reference's non-streaming branch (non-ACGT bytes count as 'A')."""
        bases, offsets = _as_batch(bases, offsets)
        n = len(offsets) - 1
        lens = np.diff(offsets).astype(np.int64)
        if (lens < 0).any():
            raise ValueError("offsets must be non-decreasing")
        if len(offsets) and int(offsets[-1]) > len(bases):
            raise ValueError("offsets index past the end of bases")
        # capacity: clean contigs give L-k+1; contigs with non-ACGT bytes can give up to L-m+1
        cap = int(np.maximum(lens - self.m + 1, 0).sum()) if out is None else len(out)
        codes = np.empty(max(cap, 1), dtype=np.uint64) if out is None else out
        code_off = np.empty(n + 1, dtype=np.uint64)
        total = C.c_uint64(0)
        fn = lib().lphb_query_stream if streaming else lib().lphb_query_nonstreaming
        _check(fn(self._h, bases.ctypes.data, offsets.ctypes.data, n,
                  codes.ctypes.data, cap, code_off.ctypes.data, C.byref(total)))
        return codes[: total.value], code_off

    def query_batch_runs(self, bases, offsets, out: np.ndarray | None = None):
        """lphb_query_stream_runs: the codes of a batch in run-length form (RUN_DTYPE records, about
        2 bytes per k-mer over PCIe instead of 8).  Returns (runs, code_offsets, n_codes)."""
        bases, offsets = _as_batch(bases, offsets)
        n = len(offsets) - 1
        lens = np.diff(offsets).astype(np.int64)
        cap = int(np.maximum(lens - self.m + 1, 0).sum()) if out is None else len(out)
        runs = np.empty(max(cap, 1), dtype=RUN_DTYPE) if out is None else out
        code_off = np.empty(n + 1, dtype=np.uint64)
        n_runs, total = C.c_uint64(0), C.c_uint64(0)
        _check(lib().lphb_query_stream_runs(self._h, bases.ctypes.data, offsets.ctypes.data, n, runs.ctypes.data,
                                            cap, C.byref(n_runs), code_off.ctypes.data, C.byref(total)))
        return runs[: n_runs.value], code_off, total.value

    def query_device(self, d_bases: int, d_offsets: int, h_offsets: np.ndarray, d_codes: int,
                     capacity: int, d_code_offsets: int, d_status: int, stream: int = 0) -> None:
        """Device-resident asynchronous variant; arguments are raw device pointers (ints), e.g.
        torch_tensor.data_ptr(), and a cudaStream_t handle."""
        h_offsets = np.ascontiguousarray(h_offsets, dtype=np.uint64)
        _check(lib().lphb_query_stream_device(self._h, d_bases, d_offsets, h_offsets.ctypes.data,
                                              len(h_offsets) - 1, d_codes, capacity, d_code_offsets,
                                              d_status, stream))


RUN_DTYPE = np.dtype([("first", "<u8"), ("n", "<i4")])  # 12 bytes packed: lphb_query_stream_runs records


def expand_runs(runs, threads: int = 1) -> np.ndarray:
    """lphb_expand_runs: run records -> the uint64 codes lphb_query_stream returns (host only)."""
    runs = np.ascontiguousarray(runs, dtype=RUN_DTYPE)
    total = int(np.abs(runs["n"].astype(np.int64)).sum())
    codes = np.empty(max(total, 1), dtype=np.uint64)
    n = C.c_uint64(0)
    _check(lib().lphb_expand_runs(runs.ctypes.data, len(runs), codes.ctypes.data, total, C.byref(n), threads))
    return codes[: n.value]


# ---- build-side scan (module-level like the reference's free functions in namespace minimizer) --

def scan_superkmers(bases, offsets, k: int, m: int, seed: int = 42, mm_count: int = 0, device: int = 0):
    """minimizer::from_string over a batch (/root/reference/include/minimizer.hpp:11-170).
    Returns (records[RECORD_DTYPE] in scan order, n_kmers, mm_count_out)."""
    bases, offsets = _as_batch(bases, offsets)
    n = len(offsets) - 1
    lens = np.diff(offsets).astype(np.int64)
    cap = int(np.maximum(lens - k + 1, 0).sum()) + n + 1
    rec = np.empty(max(cap, 1), dtype=RECORD_DTYPE)
    mm = C.c_uint64(mm_count)
    nrec, nk = C.c_uint64(0), C.c_uint64(0)
    _check(lib().lphb_scan_superkmers(device, k, m, seed, bases.ctypes.data, offsets.ctypes.data, n,
                                      C.byref(mm), rec.ctypes.data, cap, C.byref(nrec), C.byref(nk)))
    return rec[: nrec.value], nk.value, mm.value


def scan_superkmers_device(d_bases: int, d_offsets: int, h_offsets, k: int, m: int, seed: int = 42,
                           mm_count: int = 0, device: int = 0, fetch: bool = True):
    """lphb_scan_superkmers_device: bases / offsets already on the device (raw pointers).  Returns
    (records or None, n_records, n_kmers, mm_count_out, kernel_ms)."""
    h_offsets = np.ascontiguousarray(h_offsets, dtype=np.uint64)
    mm = C.c_uint64(mm_count)
    ptr = C.c_void_p()
    nrec, nk, ms = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    _check(lib().lphb_scan_superkmers_device(device, k, m, seed, d_bases, d_offsets, h_offsets.ctypes.data,
                                             len(h_offsets) - 1, C.byref(mm), C.byref(ptr), C.byref(nrec), C.byref(nk),
                                             C.byref(ms)))
    rec = None
    if fetch:
        rec = np.empty(nrec.value, dtype=RECORD_DTYPE)
        if nrec.value:
            _check(lib().lphb_copy_to_host(device, rec.ctypes.data, ptr, nrec.value * 18))
    return rec, nrec.value, nk.value, mm.value, ms.value


TRIPLET_DTYPE = np.dtype([("itself", "<u8"), ("p1", "u1"), ("size", "u1")])  # mm_triplet_t


def classify(records, device: int = 0):
    """Sort by minimizer + minimizer::classify (/root/reference/src/minimizer.cpp:5-50).
    Returns (triplets[TRIPLET_DTYPE] in ascending minimizer order, ascending colliding ids)."""
    records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    n = len(records)
    trip = np.empty(max(n, 1), dtype=TRIPLET_DTYPE)
    ids = np.empty(max(n, 1), dtype=np.uint64)
    nt, ni = C.c_uint64(0), C.c_uint64(0)
    _check(lib().lphb_classify(device, records.ctypes.data, n, trip.ctypes.data, n, C.byref(nt),
                               ids.ctypes.data, n, C.byref(ni)))
    return trip[: nt.value].copy(), ids[: ni.value].copy()


def scan_classify(bases, offsets, k: int, m: int, seed: int = 42, mm_count: int = 0, device: int = 0):
    """from_string over the batch + sort + classify with the records kept on the device.
    Returns (triplets, colliding ids, n_kmers, mm_count_out)."""
    bases, offsets = _as_batch(bases, offsets)
    n = len(offsets) - 1
    cap = int(np.maximum(np.diff(offsets).astype(np.int64) - k + 1, 0).sum()) + 1
    trip = np.empty(cap, dtype=TRIPLET_DTYPE)
    ids = np.empty(cap, dtype=np.uint64)
    mm = C.c_uint64(mm_count)
    nt, ni, nk = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    _check(lib().lphb_scan_classify(device, k, m, seed, bases.ctypes.data, offsets.ctypes.data, n, C.byref(mm),
                                    trip.ctypes.data, cap, C.byref(nt), ids.ctypes.data, cap, C.byref(ni),
                                    C.byref(nk)))
    return trip[: nt.value].copy(), ids[: ni.value].copy(), nk.value, mm.value


def colliding_kmers(bases, offsets, k: int, m: int, ids, seed: int = 42, kmer_bits: int = 64,
                    mm_count: int = 0, device: int = 0) -> np.ndarray:
    """minimizer::get_colliding_kmers over a batch (/root/reference/include/minimizer.hpp:172-319).
    Returns an (n, kmer_bits // 64) uint64 array, column 0 = low word."""
    bases, offsets = _as_batch(bases, offsets)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    n = len(offsets) - 1
    words = kmer_bits // 64
    cap = int(np.maximum(np.diff(offsets).astype(np.int64) - k + 1, 0).sum())
    out = np.empty((max(cap, 1), words), dtype=np.uint64)
    mm = C.c_uint64(mm_count)
    nk = C.c_uint64(0)
    _check(lib().lphb_colliding_kmers(device, k, m, seed, bases.ctypes.data, offsets.ctypes.data, n,
                                      C.byref(mm), ids.ctypes.data, len(ids), kmer_bits,
                                      out.ctypes.data, cap, C.byref(nk)))
    return out[: nk.value]


def build_inverted_index(k: int, m: int, minimizer_order: bytes, triplets, device: int = 0, capacity: int | None = None):
    """build-p Part 3 (re-key by minimizer_order + mphf::build_inverted_index, /root/reference/src/
    partitioned_mphf.cpp:92-106, 163-268).  minimizer_order = serialized single_phf; triplets = TRIPLET_DTYPE array.
    Returns (InvertedIndex, body bytes = image of wtree + image of sizes_and_positions)."""
    trip = np.ascontiguousarray(triplets, dtype=TRIPLET_DTYPE)
    phf = np.frombuffer(minimizer_order, dtype=np.uint8)
    cap = lib().lphb_inverted_index_bound(len(trip)) if capacity is None else capacity
    out = np.empty(max(cap, 1), dtype=np.uint8)
    info = InvertedIndex()
    nb = C.c_uint64(0)
    _check(lib().lphb_build_inverted_index(device, k, m, phf.ctypes.data, len(phf), trip.ctypes.data, len(trip),
                                           out.ctypes.data, cap, C.byref(nb), C.byref(info)))
    return info, out[: nb.value].tobytes()


def lph_sections(image: bytes, kmer_bits: int = 64, alt: bool = False) -> list[int]:
    """Start of minimizer_order, of the wavelet tree (alt: positions), of sizes_and_positions (alt: sizes), of
    fallback_kmer_order, and the end of a serialized index.  Host only."""
    buf = np.frombuffer(image, dtype=np.uint8)
    sec = (C.c_uint64 * 5)()
    _check(lib().lphb_lph_sections(buf.ctypes.data, len(buf), kmer_bits, int(alt), sec))
    return [int(v) for v in sec]


def lph_assemble(k: int, m: int, mm_seed: int, nkmers: int, distinct_minimizers: int, index: InvertedIndex,
                 minimizer_order: bytes, index_body: bytes, fallback_kmer_order: bytes) -> bytes:
    """A complete `.lph` image (what essentials::save writes, /root/reference/src/build.cpp:52) from its parts.  Host only."""
    mo = np.frombuffer(minimizer_order, dtype=np.uint8)
    body = np.frombuffer(index_body, dtype=np.uint8)
    fb = np.frombuffer(fallback_kmer_order, dtype=np.uint8)
    cap = 58 + len(mo) + len(body) + len(fb)
    out = np.empty(cap, dtype=np.uint8)
    nb = C.c_uint64(0)
    _check(lib().lphb_lph_assemble(k, m, mm_seed, nkmers, distinct_minimizers, C.byref(index), mo.ctypes.data, len(mo),
                                   body.ctypes.data, len(body), fb.ctypes.data, len(fb), out.ctypes.data, cap, C.byref(nb)))
    return out[: nb.value].tobytes()


def build_inverted_index_alt(minimizer_order: bytes, triplets, device: int = 0):
    """build-u Part 3 (mphf_alt: /root/reference/src/unpartitioned_mphf.cpp:78-96, 152-169).
    Returns (InvertedIndexAlt, body bytes = image of positions + image of sizes)."""
    trip = np.ascontiguousarray(triplets, dtype=TRIPLET_DTYPE)
    phf = np.frombuffer(minimizer_order, dtype=np.uint8)
    cap = lib().lphb_inverted_index_bound(len(trip))
    out = np.empty(max(cap, 1), dtype=np.uint8)
    info = InvertedIndexAlt()
    nb = C.c_uint64(0)
    _check(lib().lphb_build_inverted_index_alt(device, phf.ctypes.data, len(phf), trip.ctypes.data, len(trip),
                                               out.ctypes.data, cap, C.byref(nb), C.byref(info)))
    return info, out[: nb.value].tobytes()


def lph_assemble_alt(k: int, m: int, mm_seed: int, nkmers: int, distinct_minimizers: int, index: InvertedIndexAlt,
                     minimizer_order: bytes, index_body: bytes, fallback_kmer_order: bytes) -> bytes:
    """A complete serialized lphash::mphf_alt (what build-u saves) from its parts.  Host only."""
    mo = np.frombuffer(minimizer_order, dtype=np.uint8)
    body = np.frombuffer(index_body, dtype=np.uint8)
    fb = np.frombuffer(fallback_kmer_order, dtype=np.uint8)
    cap = 34 + len(mo) + len(body) + len(fb)
    out = np.empty(cap, dtype=np.uint8)
    nb = C.c_uint64(0)
    _check(lib().lphb_lph_assemble_alt(k, m, mm_seed, nkmers, distinct_minimizers, C.byref(index), mo.ctypes.data, len(mo),
                                       body.ctypes.data, len(body), fb.ctypes.data, len(fb), out.ctypes.data, cap,
                                       C.byref(nb)))
    return out[: nb.value].tobytes()
