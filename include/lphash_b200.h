/*
 * lphash_b200 — C ABI of the B200-native LPHash hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions.
 * The reference (jermp/lphash) has no FFI of its own; its seam for this path is a C++ member
 * function, so each entry point below cites the reference interface it replaces
 * (paths relative to the reference tree; "pthash/" = external/pthash/).  INTEGRATION.md shows
 * the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns LPHB_OK (0) or a negative LPHB_E_* code; lphb_last_error() gives the
 *     message of the last failure on the calling thread;
 *   - the caller owns all host buffers; the library owns device memory behind opaque handles;
 *   - one handle = one `.lph` image resident on one GPU; calls on one handle are serialised by
 *     the caller, different handles (different GPUs) may be driven from different host threads;
 *   - "batch" = ASCII bases of n_contigs contigs concatenated without separators + offsets
 *     (n_contigs + 1 entries, offsets[0] need not be 0 but must index `bases`); this is what
 *     kseq hands the reference one record at a time (src/query.cpp:51-52);
 *   - there is NO CPU fallback: every call that computes needs a CUDA device and fails with
 *     LPHB_E_CUDA otherwise.
 */
#ifndef LPHASH_B200_H
#define LPHASH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPHB_OK 0
#define LPHB_E_ARG (-1)      /* bad argument (null pointer, k/m out of range, ...)            */
#define LPHB_E_IO (-2)       /* cannot open / read the .lph file                              */
#define LPHB_E_FORMAT (-3)   /* not a partitioned-LP-MPHF image (truncated, trailing bytes...) */
#define LPHB_E_CUDA (-4)     /* CUDA error or no device                                       */
#define LPHB_E_CAPACITY (-5) /* output buffer too small; required size reported where possible */
#define LPHB_E_NOMEM (-6)

typedef struct lphb_mphf lphb_mphf; /* device image of one lphash::mphf on one GPU */

/* Header fields of the serialized lphash::mphf (include/partitioned_mphf.hpp:204-211, :42-48). */
typedef struct lphb_info {
    uint32_t k, m;
    uint32_t kmer_bits; /* 64 or 128: the kmer_t the file was built with (compile_constants.tpd) */
    int32_t device;
    uint64_t mm_seed;
    uint64_t nkmers;
    uint64_t distinct_minimizers;
    uint64_t n_maximal;
    uint64_t right_coll_sizes_start;
    uint64_t none_sizes_start;
    uint64_t none_pos_start;
    uint64_t fallback_keys; /* k-mers handled by fallback_kmer_order */
    uint64_t file_bytes;    /* size of the .lph image parsed                                    */
    uint64_t device_bytes;  /* bytes of HBM the device image occupies                           */
    double load_host_ms;    /* load time on the host: structure walk of the image (+ CUDA start-up if this
                               is the process's first CUDA call; + the decode with LPHB_HOST_DECODE=1) */
    double load_h2d_ms;     /* load time on the device: allocation, upload, decode kernels          */
} lphb_info;

const char* lphb_last_error(void);
const char* lphb_version(void);
int lphb_device_count(int* count);

/* ---- loading: replaces essentials::load(hf, file) (src/query.cpp:37) ----------------------
 * Parses the byte-exact essentials visitor image of lphash::mphf
 * (include/partitioned_mphf.hpp:204-219) and builds a flat device image from it.  The host only walks the
 * structure (sizes, ranges: LPHB_E_FORMAT before any CUDA call for a malformed file); the decoding -
 * compact_vector::access, Elias-Fano access, rs_bit_vector::rank, quartet_wtree::rank_of, one word per bucket -
 * runs on the GPU (lphash_b200/csrc/image_decode.cu) and reports what only decoding reveals as
 * LPHB_E_FORMAT too.  kmer_bits selects the fallback hash flavour of include/constants.hpp:56-70 (the
 * file does not record it).                                                                      */
int lphb_mphf_load_file(const char* path, int kmer_bits, int device, lphb_mphf** out);
int lphb_mphf_load_memory(const void* image, uint64_t nbytes, int kmer_bits, int device,
                          lphb_mphf** out);
/* The same for the UNPARTITIONED variant, lphash::mphf_alt (`build-u` / `query-u`: include/
 * unpartitioned_mphf.hpp:198-210 is the file format, src/unpartitioned_mphf.cpp:191-206 the probe).  The
 * handle then serves every query entry point below exactly like a partitioned one: mphf_alt's
 * operator() (include/unpartitioned_mphf.hpp:72-192) is the same streaming loop over another probe.
 * (lphb_info: n_maximal and the *_start fields are 0.)                                              */
int lphb_mphf_alt_load_file(const char* path, int kmer_bits, int device, lphb_mphf** out);
int lphb_mphf_alt_load_memory(const void* image, uint64_t nbytes, int kmer_bits, int device,
                              lphb_mphf** out);
int lphb_mphf_free(lphb_mphf* f);
int lphb_mphf_info(const lphb_mphf* f, lphb_info* info);

/* ---- query-p hot call ----------------------------------------------------------------------
 * Replaces `hf(seq->seq.s, seq->seq.l, true)` = lphash::mphf::operator()(contig, len,
 * streaming=true) (include/partitioned_mphf.hpp:73-184; caller src/query.cpp:52) for a whole
 * batch of contigs.  HOST buffers in, HOST buffers out (H2D / D2H copies happen inside).
 *   codes         receives the hash codes of all contigs, concatenated in contig order
 *   code_offsets  (n_contigs + 1) receives where each contig's codes start; contig c produced
 *                 code_offsets[c+1] - code_offsets[c] codes: L-k+1 for an ACGT-only contig of
 *                 length L >= k, 0 if L < k; contigs with other bytes follow the reference's
 *                 streaming behaviour exactly (its spurious entries included)
 *   n_codes       total number of codes (also set when returning LPHB_E_CAPACITY)              */
int lphb_query_stream(lphb_mphf* f, const char* bases, const uint64_t* offsets,
                      uint64_t n_contigs, uint64_t* codes, uint64_t codes_capacity,
                      uint64_t* code_offsets, uint64_t* n_codes);

/* The reference's NON-streaming branch, hf(seq, len, false) (include/partitioned_mphf.hpp:185-195;
 * pass 2 of query-p, src/query.cpp:72, and `--check`, include/mphf_utils.hpp:56,85): the stateless
 * evaluation of every window of k bytes, where a non-ACGT byte counts as 'A'
 * (include/mphf_utils.hpp:108).  Same arguments as lphb_query_stream; contig c always yields
 * max(0, L-k+1) codes (the reference itself is undefined for L < k).  On ACGT-only input the two
 * branches return the same codes.                                                  */
int lphb_query_nonstreaming(lphb_mphf* f, const char* bases, const uint64_t* offsets,
                            uint64_t n_contigs, uint64_t* codes, uint64_t codes_capacity,
                            uint64_t* code_offsets, uint64_t* n_codes);

/* Same call with the codes returned in RUN-LENGTH form: consecutive k-mers of a super-k-mer get
 * consecutive codes (include/partitioned_mphf.hpp:131-145: local_rank +-1 per k-mer), so the vector
 * the reference returns is a sequence of arithmetic runs.  `runs` receives packed 12-byte records
 * {uint64_t first; int32_t n} (little-endian; n > 0: first, first+1, ... n codes; n < 0: first,
 * first-1, ... |n| codes), in stream order and never across a contig start: about 2 bytes per k-mer
 * over PCIe instead of 8.  lphb_expand_runs rebuilds the identical uint64_t vector on the host.
 * Works for any input (non-members, colliding minimizers, non-ACGT contigs: their runs are just
 * shorter).  A batch must hold fewer than 2^31 k-mers.  n_runs is also set when returning
 * LPHB_E_CAPACITY (which is also what runs == NULL returns after reporting n_runs / n_codes /
 * code_offsets).                                                                                  */
int lphb_query_stream_runs(lphb_mphf* f, const char* bases, const uint64_t* offsets,
                           uint64_t n_contigs, void* runs, uint64_t runs_capacity, uint64_t* n_runs,
                           uint64_t* code_offsets, uint64_t* n_codes);

/* Host-only decoding of run records into the codes of lphb_query_stream (same order, same values);
 * `threads` host threads share the work (<= 1: the calling thread).  n_codes is also set when
 * returning LPHB_E_CAPACITY.                                                                       */
int lphb_expand_runs(const void* runs, uint64_t n_runs, uint64_t* codes, uint64_t codes_capacity,
                     uint64_t* n_codes, int threads);

/* Device-resident variant: d_* are device pointers on the handle's GPU, `stream` is a
 * cudaStream_t (NULL = default stream).  Asynchronous.  h_offsets is the host copy of the same
 * offsets (used only for sizes).  Evaluates the stateless definition (one code per window of k
 * ACGT bases); d_status[0] = number of codes laid out (sum of max(0, L-k+1)),
 * d_status[1] = number of contigs containing a non-ACGT byte.  For such a contig the code slot of
 * every k-mer window that contains a non-ACGT byte is UNSPECIFIED (the slots of its other k-mers hold
 * their codes); lphb_mphf_dirty_flags names those contigs, and lphb_query_stream applies the
 * reference's exact streaming behaviour to them - this call does not.
 * Ordering: a handle owns one device workspace, so consecutive calls on one handle are ordered on
 * the device (each call makes its stream wait for the previous call's kernels), on whatever
 * streams they are issued; they never overlap each other.  Overlap comes from several handles or
 * from the copies of the caller's other streams.                                                  */
int lphb_query_stream_device(lphb_mphf* f, const char* d_bases, const uint64_t* d_offsets,
                             const uint64_t* h_offsets, uint64_t n_contigs, uint64_t* d_codes,
                             uint64_t codes_capacity, uint64_t* d_code_offsets,
                             uint64_t* d_status, void* stream);

/* Diagnostic: the flat device image behind the handle (layout: lphash_b200/csrc/device_image.h), e.g. to compare
 * the GPU-side decode of a load with the host-side one (LPHB_HOST_DECODE=1).  Read it with lphb_copy_to_host.  */
int lphb_mphf_device_image(const lphb_mphf* f, const void** d_image, uint64_t* nbytes);

/* Per-contig flags of the last query call on the handle: *d_flags = DEVICE pointer to *n_contigs
 * bytes, nonzero where the contig contains a non-ACGT byte; valid (in stream order) until the next
 * query call on the handle.                                                                        */
int lphb_mphf_dirty_flags(const lphb_mphf* f, const uint8_t** d_flags, uint64_t* n_contigs);

/* ---- build-p Part 1: minimizer / super-k-mer scan -------------------------------------------
 * Replaces the loop over minimizer::from_string (include/minimizer.hpp:11-170; caller
 * src/partitioned_mphf.cpp:70-77) for a batch of contigs.  records receives packed 18-byte
 * mm_record_t {u64 itself, u64 id, u8 p1, u8 size} (include/constants.hpp:26-33) in scan order;
 * mm_count is the running global m-mer ordinal (in: value before the batch, out: after).
 * Bytes other than ACGT/acgt (U/u) are handled as the reference's loop does (include/minimizer.hpp:
 * 138-151): the open super-k-mer is flushed at the byte and the window restarts, m-mer ordinals advance
 * over valid runs only, and a run of exactly k valid bases that is followed by such a byte counts its
 * k-mer in *n_kmers without ever emitting a record for it.  (Clean batches - the unitig sets lphash is
 * built from - take one kernel; a batch with such bytes is cut at them on the device and scanned again.)
 * A batch must hold fewer than 2^32 k-mers and span fewer than 2^32 bases.                         */
int lphb_scan_superkmers(int device, uint32_t k, uint32_t m, uint64_t seed, const char* bases,
                         const uint64_t* offsets, uint64_t n_contigs, uint64_t* mm_count,
                         void* records, uint64_t records_capacity, uint64_t* n_records,
                         uint64_t* n_kmers);

/* Device-resident variant: d_bases / d_offsets are device pointers on `device` (d_bases indexed by
 * the offsets), h_offsets the host copy of the offsets.  Synchronous.  *d_records receives a DEVICE
 * pointer to the packed records (scan order), owned by the library and valid until the next scan
 * call on the device or lphb_scan_release; *kernel_ms (optional) the CUDA-event time of the
 * record-producing kernels.                                                                        */
int lphb_scan_superkmers_device(int device, uint32_t k, uint32_t m, uint64_t seed, const char* d_bases,
                                const uint64_t* d_offsets, const uint64_t* h_offsets, uint64_t n_contigs,
                                uint64_t* mm_count, const void** d_records, uint64_t* n_records,
                                uint64_t* n_kmers, double* kernel_ms);

/* The two build-side entry points keep their device workspace per device between calls (a caller
 * streaming batches does not pay allocation on every batch); this frees it.                      */
int lphb_scan_release(int device);

/* ---- build-p Part 2: sort by minimizer + classify ----------------------------------------------
 * Replaces the sort of the mm_record_t stream by minimizer (external_memory_vector<mm_record_t>,
 * src/partitioned_mphf.cpp:62-66) and minimizer::classify (src/minimizer.cpp:5-50; caller
 * src/partitioned_mphf.cpp:86).  records = n_records packed 18-byte mm_record_t in any order.
 * triplets receives packed 10-byte mm_triplet_t {u64 itself, u8 p1, u8 size}
 * (include/constants.hpp:35-41), one per distinct minimizer in ascending minimizer order -
 * {itself, p1, size} for a minimizer seen once, {itself, 0, 0} for one seen several times - i.e. the
 * key stream PTHash consumes; ids receives the ids of all occurrences of the latter, ascending.
 * n_triplets / n_ids are also set when returning LPHB_E_CAPACITY.                                 */
int lphb_classify(int device, const void* records, uint64_t n_records, void* triplets,
                  uint64_t triplets_capacity, uint64_t* n_triplets, uint64_t* ids,
                  uint64_t ids_capacity, uint64_t* n_ids);

/* Parts 1 + 2 in one call: lphb_scan_superkmers followed by lphb_classify with the record stream
 * kept on the device (what mphf::build does between src/partitioned_mphf.cpp:62 and :86).        */
int lphb_scan_classify(int device, uint32_t k, uint32_t m, uint64_t seed, const char* bases,
                       const uint64_t* offsets, uint64_t n_contigs, uint64_t* mm_count, void* triplets,
                       uint64_t triplets_capacity, uint64_t* n_triplets, uint64_t* ids,
                       uint64_t ids_capacity, uint64_t* n_ids, uint64_t* n_kmers);

/* ---- build-p Part 3: the inverted index --------------------------------------------------------
 * Replaces the loop that re-keys every triplet by minimizer_order(itself) and sorts the result
 * (src/partitioned_mphf.cpp:92-106) and mphf::build_inverted_index (src/partitioned_mphf.cpp:163-268):
 * type of every distinct minimizer, the quartet wavelet tree with its rank directories
 * (src/quartet_wtree.cpp:13-53, include/rs_bit_vector.hpp:120-156), the four size / position lists and
 * the Elias-Fano encoding of their prefix sums (include/ef_sequence.hpp:36-75) with its darray1 select
 * index (pthash/include/encoders/darray.hpp:13-48).
 * minimizer_order = the serialized pthash::single_phf built on the distinct minimizers (the bytes
 * essentials::save writes for it; PTHash construction itself stays with the caller); triplets = the
 * packed 10-byte mm_triplet_t of lphb_classify, one per distinct minimizer, in any order.
 * out receives, byte for byte as in the `.lph` file, the image of `wtree` followed by the image of
 * `sizes_and_positions` (include/partitioned_mphf.hpp:213-214); info the four counters stored in front
 * of them.  lphb_inverted_index_bound(n) is a capacity that always suffices; *out_bytes is also set
 * when returning LPHB_E_CAPACITY.  LPHB_E_ARG if minimizer_order is not a bijection of the triplets'
 * minimizers onto [0, n_triplets).                                                                   */
typedef struct lphb_inverted_index {
    uint64_t n_maximal, right_coll_sizes_start, none_sizes_start, none_pos_start; /* include/partitioned_mphf.hpp:208-211 */
    uint64_t colliding_minimizers; /* triplets of size 0 */
    uint64_t universe;             /* last prefix sum (src/partitioned_mphf.cpp:267) */
    uint64_t wtree_bytes, ef_bytes; /* the two images inside out */
    double device_ms;              /* CUDA-event time of the kernels, copies excluded */
} lphb_inverted_index;
uint64_t lphb_inverted_index_bound(uint64_t n_triplets);
int lphb_build_inverted_index(int device, uint32_t k, uint32_t m, const void* minimizer_order,
                              uint64_t minimizer_order_bytes, const void* triplets, uint64_t n_triplets,
                              void* out, uint64_t out_capacity, uint64_t* out_bytes,
                              lphb_inverted_index* info);

/* The same for the unpartitioned variant (build-u, lphash::mphf_alt; src/unpartitioned_mphf.cpp:78-96 re-key and
 * :152-169 build_pos_index / build_size_index): out receives the image of `positions` (Elias-Fano of the prefix sums
 * of p1 in minimizer_order order) followed by the image of `sizes` (include/unpartitioned_mphf.hpp:206-207);
 * lphb_inverted_index_bound(n_triplets) suffices as capacity.                                              */
typedef struct lphb_inverted_index_alt {
    uint64_t num_kmers_in_main_index;      /* last prefix sum of the sizes (src/unpartitioned_mphf.cpp:168) */
    uint64_t positions_bytes, sizes_bytes; /* the two images inside out */
    double device_ms;
} lphb_inverted_index_alt;
int lphb_build_inverted_index_alt(int device, const void* minimizer_order, uint64_t minimizer_order_bytes,
                                  const void* triplets, uint64_t n_triplets, void* out, uint64_t out_capacity,
                                  uint64_t* out_bytes, lphb_inverted_index_alt* info);
int lphb_lph_assemble_alt(uint32_t k, uint32_t m, uint64_t mm_seed, uint64_t nkmers, uint64_t distinct_minimizers,
                          const lphb_inverted_index_alt* index, const void* minimizer_order,
                          uint64_t minimizer_order_bytes, const void* index_body, uint64_t index_body_bytes,
                          const void* fallback_kmer_order, uint64_t fallback_bytes, void* out,
                          uint64_t out_capacity, uint64_t* out_bytes);

/* ---- the `.lph` writer (host only) ------------------------------------------------------------
 * lphb_lph_assemble lays out a complete serialized lphash::mphf in the visitor order of
 * include/partitioned_mphf.hpp:204-219 - what essentials::save(hf, file) writes (src/build.cpp:52) -
 * from the header fields, the two serialized single_phf objects (built by PTHash on the caller's side)
 * and the index body of lphb_build_inverted_index.  The result loads with the reference's
 * essentials::load and with lphb_mphf_load_memory.
 * lphb_lph_sections reports where the parts of an existing image start: minimizer_order, the wavelet
 * tree (alt: positions), sizes_and_positions (alt: sizes), fallback_kmer_order, end of the image.   */
int lphb_lph_assemble(uint32_t k, uint32_t m, uint64_t mm_seed, uint64_t nkmers, uint64_t distinct_minimizers,
                      const lphb_inverted_index* index, const void* minimizer_order,
                      uint64_t minimizer_order_bytes, const void* index_body, uint64_t index_body_bytes,
                      const void* fallback_kmer_order, uint64_t fallback_bytes, void* out,
                      uint64_t out_capacity, uint64_t* out_bytes);
int lphb_lph_sections(const void* image, uint64_t nbytes, int kmer_bits, int alt, uint64_t sections[5]);

/* ---- build-p Part 4: k-mers of colliding minimizers -----------------------------------------
 * Replaces the loop over minimizer::get_colliding_kmers (include/minimizer.hpp:172-319; caller
 * src/partitioned_mphf.cpp:120-129).  ids = ascending minimizer-occurrence ids (classify's second
 * output).  kmers receives kmer_bits/8 little-endian bytes per k-mer, in scan order.            */
int lphb_colliding_kmers(int device, uint32_t k, uint32_t m, uint64_t seed, const char* bases,
                         const uint64_t* offsets, uint64_t n_contigs, uint64_t* mm_count,
                         const uint64_t* ids, uint64_t n_ids, int kmer_bits, void* kmers,
                         uint64_t kmers_capacity, uint64_t* n_kmers);

/* ---- pinned host memory (optional; makes the copies inside lphb_query_stream asynchronous) -- */
int lphb_host_alloc(void** ptr, uint64_t nbytes);
/* synchronous copy of a library-owned device buffer (e.g. *d_records of lphb_scan_superkmers_device) */
int lphb_copy_to_host(int device, void* dst, const void* d_src, uint64_t nbytes);
int lphb_host_free(void* ptr);

/* Counters for the last lphb_query_stream* call on the handle (kernel launches, device ms). */
typedef struct lphb_stats {
    uint64_t kernel_launches;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t dirty_contigs;
    double kernel_ms; /* mean CUDA-event time of the main query kernel over the calls since the last
                         lphb_mphf_stats (most recent 128) */
} lphb_stats;
int lphb_mphf_stats(const lphb_mphf* f, lphb_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* LPHASH_B200_H */
