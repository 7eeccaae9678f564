// Launchers of the query kernels (query_generic.cu, query_tiled.cu).  All pointers are device
// pointers; all launches are asynchronous on `stream`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_image.h"

namespace lphb {

// Per-batch device-side description of the contig layout.
struct DevBatch {
    const char* bases;          // concatenated ASCII
    const uint64_t* offsets;    // n_contigs + 1, indexes `bases`
    const uint64_t* code_off;   // n_contigs + 1: exclusive scan of max(0, L - k + 1)
    uint64_t n_contigs;
    uint64_t first_base;        // offsets[0]
    uint64_t end_base;          // offsets[n_contigs]
    uint64_t* codes;            // output
    uint8_t* dirty;             // n_contigs flags: contig contains a non-ACGT byte
    unsigned long long* status; // [0] unused here, [1] += number of dirty contigs
    void* tile_ws;              // scratch for the tiled kernel (query_tiled_ws_bytes), may be null
    uint64_t tile_ws_bytes;
};
uint64_t query_tiled_ws_bytes(uint64_t span_bases);

// code_off[c] = sum_{c' < c} max(0, L_c' - k + 1); status[0] = total.  One launch, any n.
void launch_code_offsets(const uint64_t* d_offsets, uint64_t n_contigs, uint32_t k,
                         uint64_t* d_code_off, unsigned long long* d_status, void* d_tmp,
                         uint64_t tmp_bytes, cudaStream_t stream);
uint64_t code_offsets_tmp_bytes(uint64_t n_contigs);

// Generic kernel: any (k, m); one thread per k-mer, everything recomputed per k-mer.
void launch_query_generic(DevImage const& img, DevBatch const& b, cudaStream_t stream);

// Tiled kernel for the (k, m) pairs it is instantiated for; returns false if (k, m) is not one
// of them (the caller then uses the generic kernel).
bool launch_query_tiled(DevImage const& img, DevBatch const& b, cudaStream_t stream);

// Build-side form of the tiled kernel (minimizer::from_string): writes the packed 18-byte mm_record_t
// stream of the batch in scan order to d_records (capacity: one record per k-mer start), the stream
// position (relative to b.first_base) of every record's first k-mer to d_start_pos, and the number of
// records to *d_n_records; flags contigs with non-ACGT bytes in b.dirty (their records are then
// meaningless).  d_id_base[c] = m-mer ordinal of contig c's first m-mer.  b.codes / b.code_off are not
// used.  Returns false if (k, m) is not instantiated (the caller then uses the generic scan kernels).
bool launch_scan_records_tiled(uint32_t k, uint32_t m, uint64_t seed, DevBatch const& b, const uint64_t* d_id_base,
                               uint8_t* d_records, uint32_t* d_start_pos, unsigned long long* d_n_records,
                               cudaStream_t stream);
bool scan_tiled_available(uint32_t k, uint32_t m);

// Contigs with non-ACGT bytes, run-parallel (quirk_kernels.cu).  All asynchronous on `stream`.
void launch_gather_contigs(const char* bases, const uint64_t* offsets, const uint64_t* list, const uint64_t* dst_off,
                           uint64_t n_list, uint64_t total_bytes, char* dst, cudaStream_t stream);
uint64_t quirk_tmp_bytes(uint64_t n);
void launch_find_runs(const char* d_bases, uint64_t n, const uint64_t* d_starts, uint64_t n_list, uint8_t* d_flag,
                      uint64_t* d_voff, unsigned long long* d_n_v, uint32_t* d_vstart, void* d_tmp, uint64_t tmp_bytes,
                      cudaStream_t stream);
void launch_quirk_finish(DevImage const& img, const char* d_bases, const uint64_t* d_voff, uint64_t n_v_host,
                         const unsigned long long* d_n_v, const uint32_t* d_vstart, uint64_t n_list,
                         const uint64_t* d_vcode_off, const uint64_t* d_vcodes, uint64_t* d_spur, uint32_t* d_spur_cnt,
                         uint32_t w_cap, uint64_t* d_out_off, uint64_t* d_out, uint64_t* d_q_off, uint64_t* d_counts,
                         void* d_tmp, uint64_t tmp_bytes, cudaStream_t stream);
void launch_quirk_emit(const uint64_t* d_vcode_off, const uint64_t* d_vcodes, const uint64_t* d_spur,
                       const uint32_t* d_spur_cnt, uint32_t w_cap, uint64_t n_v_host, const unsigned long long* d_n_v,
                       const uint64_t* d_out_off, uint64_t total_hint, uint64_t* d_out, cudaStream_t stream);

// Build side (minimizer::from_string over contigs with non-ACGT bytes): launch_find_runs over the whole batch
// (d_starts = contig starts relative to the batch, launch_shift_starts), then per run r the number of pieces the
// scan kernels get it in (1 for a run they must scan; invalid runs and runs of exactly k bases followed by an
// invalid byte are cut into pieces of k-1 bytes, which hold no k-mer), its m-mer ordinals and k-mers - exclusive
// sums in place, entry n_v = totals - and finally the piece offsets (n_pieces + 1, relative to the batch) with the
// m-mer ordinal of every piece's first m-mer.
void launch_shift_starts(const uint64_t* d_offsets, uint64_t n_contigs, uint64_t first, uint64_t* d_starts, cudaStream_t stream);
void launch_build_run_counts(const char* d_bases, const uint64_t* d_voff, const unsigned long long* d_n_v,
                             const uint32_t* d_vstart, uint64_t n_list, uint64_t n_v_host, uint32_t k, uint32_t m,
                             uint64_t* d_pieces, uint64_t* d_ids, uint64_t* d_kmers, void* d_tmp, uint64_t tmp_bytes,
                             cudaStream_t stream);
void launch_build_pieces(const uint64_t* d_voff, const unsigned long long* d_n_v, uint64_t n_v_host, uint32_t k,
                         const uint64_t* d_pieces, const uint64_t* d_ids, uint64_t mm_count_in, uint64_t n, uint64_t* d_poff,
                         uint64_t* d_pid, cudaStream_t stream);

// dst[dst_off[c] .. dst_off[c+1]) = src_c[src_off[c] ..) where src_c = from_b[c] ? src_b : src_a (`total` = dst_off[n_contigs])
void launch_assemble(uint64_t* dst, const uint64_t* dst_off, const uint64_t* src_a,
                     const uint64_t* src_b, const uint64_t* src_off, const uint8_t* from_b,
                     uint64_t n_contigs, uint64_t total, cudaStream_t stream);

// Run-length form of a code stream on the device (runs_kernels.cu): codes[0..n), n < 2^31 -> 12-byte
// records {u64 first, i32 n} (n > 0 ascending, n < 0 descending), never across a contig start.
uint64_t runs_tmp_bytes(uint64_t n);
void launch_runs(const uint64_t* codes, uint64_t n, const uint64_t* code_off, uint64_t n_contigs, uint64_t o0,
                 uint8_t* d_start, uint8_t* d_head, uint32_t* d_rank, uint32_t* d_head_at, void* d_tmp,
                 uint64_t tmp_bytes, uint8_t* d_runs, unsigned long long* d_n_runs, cudaStream_t stream);

// non-ACGT bytes -> 'A' in place (the reference's non-streaming branch, include/mphf_utils.hpp:108)
void launch_sanitize(char* d_bases, uint64_t n, cudaStream_t stream);

// status[1] += number of set flags in dirty[0..n)
void launch_count_dirty(const uint8_t* dirty, uint64_t n, unsigned long long* status,
                        cudaStream_t stream);

}  // namespace lphb
