// Launchers of build-p Part 3 on the device (invindex_kernels.cu): re-key of the triplets by
// minimizer_order (ref src/partitioned_mphf.cpp:92-106) and mphf::build_inverted_index (ref
// src/partitioned_mphf.cpp:163-268) including the serialized forms of quartet_wtree / rs_bit_vector
// (include/rs_bit_vector.hpp:120-156), ef_sequence::encode (include/ef_sequence.hpp:36-75) and
// pthash::darray1 (pthash/include/encoders/darray.hpp:13-48, 98-120).  Device pointers, asynchronous.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_image.h"

namespace lphb {

constexpr uint32_t kInvBlock = 1024;  // cells per block of the rank pass

struct InvCounts {  // of one block of cells, or a prefix of blocks
    uint32_t left, rc, none, msb;  // LEFT, RIGHT_OR_COLLISION, NONE; msb = MAXIMAL + NONE (root bit set)
};

// cells[minimizer_order(itself)] = 0x80000000 | type | a << 2 | b << 10 for every triplet; type as in
// include/quartet_wtree.hpp:7 (LEFT 0, RIGHT_OR_COLLISION 1, MAXIMAL 2, NONE 3); a = p1 + 1 (LEFT), size
// (RIGHT_OR_COLLISION: 0 for a colliding minimizer; NONE), b = p1 (NONE).  cells must be zeroed;
// *bad counts triplets mapped outside [0, n); *colliding the triplets of size 0.
void launch_rekey(DevPhf const& phf, const uint8_t* triplets, uint64_t n, uint32_t k, uint32_t m,
                  uint32_t* cells, unsigned long long* bad, unsigned long long* colliding, cudaStream_t s);

// build-u (mphf_alt): cells[minimizer_order(itself)] = 0x80000000 | p1 | size << 8, then the two byte lists in
// order (ref src/unpartitioned_mphf.cpp:78-96, 152-168); *unset counts cells never written
void launch_rekey_alt(DevPhf const& phf, const uint8_t* triplets, uint64_t n, uint32_t* cells, unsigned long long* bad,
                      cudaStream_t s);
void launch_split_alt(const uint32_t* cells, uint64_t n, uint8_t* p1, uint8_t* size, unsigned long long* unset,
                      cudaStream_t s);

// per block of kInvBlock cells: counts by type; *unset += cells never written
void launch_cell_counts(const uint32_t* cells, uint64_t n, InvCounts* blk, unsigned long long* unset, cudaStream_t s);
// inclusive prefix over the blocks, in place
uint64_t inv_scan_tmp_bytes(uint64_t n_blocks);
void launch_counts_scan(InvCounts* blk, uint64_t n_blocks, void* tmp, uint64_t tmp_bytes, cudaStream_t s);

// wavelet-tree bits (root by ballot, the two leaves by atomicOr; all zeroed before) and the four value
// lists, concatenated: vals[0, rs) LEFT p1+1; [rs, ns) RIGHT_OR_COLLISION sizes; [ns, np) NONE sizes;
// [np, np + #NONE) NONE positions
void launch_place(const uint32_t* cells, uint64_t n, const InvCounts* incl, uint64_t rs, uint64_t ns, uint64_t np,
                  uint32_t* root, uint32_t* left_right, uint32_t* max_none, uint8_t* vals, cudaStream_t s);

// cum[i] = vals[0] + ... + vals[i]   (cumulative_iterator, include/ef_sequence.hpp:9-31)
uint64_t cum_tmp_bytes(uint64_t n);
void launch_cumulative(const uint8_t* vals, uint64_t n, uint64_t* cum, void* tmp, uint64_t tmp_bytes, cudaStream_t s);

// Elias-Fano of {0, cum[0], ..., cum[n-1]} (n_enc = n + 1 values, low width l): high bits (zeroed
// before), low bits (every word written, the spare last one included)
void launch_ef_encode(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t* high, uint64_t* low,
                      uint64_t low_words, cudaStream_t s);

// darray1 over the high bits: one block per 1024 ones.  sparse_cnt[b] = ones of block b when its span
// reaches 2^16 positions, else 0; ovf_off = exclusive sum of sparse_cnt (n_blocks + 1 entries)
uint64_t darray_tmp_bytes(uint64_t n_blocks);
void launch_darray(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t n_blocks, uint64_t* sparse_cnt,
                   uint64_t* ovf_off, void* tmp, uint64_t tmp_bytes, cudaStream_t s);
// second half, once the caller knows the overflow size and has allocated `overflow`
void launch_darray_fill(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t n_blocks, const uint64_t* sparse_cnt,
                        const uint64_t* ovf_off, int64_t* block_inventory, uint16_t* subblock_inventory,
                        uint64_t* overflow, cudaStream_t s);

// rs_bit_vector::build_indices without select hints: pairs has 2 * n_blocks + 2 entries, n_blocks =
// ceil(words / 8); `bits` is readable (zero) up to 8 * n_blocks words; pop = n_blocks + 1 scratch words
uint64_t rank_tmp_bytes(uint64_t n_blocks);
void launch_rank_pairs(const uint64_t* bits, uint64_t n_blocks, uint64_t* pop, uint64_t* pairs, void* tmp,
                       uint64_t tmp_bytes, cudaStream_t s);

}  // namespace lphb
