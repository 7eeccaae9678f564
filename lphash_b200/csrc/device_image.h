// Flat device image of one serialized lphash::mphf — the HBM-resident form of the structures the
// query probes.  POD only: passed to kernels by value (__grid_constant__).
//
// The image is NOT the serialized layout.  It holds the same information re-laid-out for the GPU
// (results are bit-identical; see lph_image.cpp for the transformations):
//   * pilots: the reference's dual<dictionary,dictionary> (pthash encoders.hpp:167-176, 268-277)
//     is two (ranks, dict) compact-vector pairs; here both halves are merged into ONE rank array
//     (u16 when the two dictionaries together have <= 65536 entries, else u32) indexing ONE table
//     that already holds default_hash64(pilot, seed) (single_phf.hpp:58): one 2-byte and one 8-byte
//     load and no second murmur per probe.
//   * rank bit-vectors (rs_bit_vector.hpp): bits and rank directory interleaved in 16-byte
//     units {u32 ones_before, u32 bits[3]} so bit + rank cost one 16-byte load, not three lines.
//   * Elias-Fano (include/ef_sequence.hpp, pthash ef_sequence.hpp): high bits, darray
//     inventories and low bits kept as in the file, word-aligned and padded.
#pragma once
#include <stdint.h>
#include <vector_types.h>  // uint4 (plain struct; usable from host-only translation units)

namespace lphb {

struct DevCompact {          // pthash::compact_vector (compact_vector.hpp:277-283), aligned
    const uint64_t* bits;    // padded with >= 1 zero word past the end
    uint64_t size;
    uint32_t width;          // 0 is legal (EF low bits): get() == 0
    uint64_t mask;
};

struct DevEF {               // {high bit_vector, darray1, low compact_vector}
    const uint64_t* high;    // padded with 1 word
    const int64_t* block_inv;      // darray.hpp: one entry per 1024 ones (<0: overflow index)
    const uint16_t* sub_inv;       // one entry per 32 ones
    const uint64_t* overflow;
    DevCompact low;
    uint64_t n;              // number of encoded values
};

// sizes_and_positions re-laid-out for one-sector lookups.  The reference stores the prefix sums
// S[0]=0, S[i+1]=S[i]+d[i] in Elias-Fano form (include/ef_sequence.hpp) and reads S[i], S[i+1];
// every d[i] is a super-k-mer size / minimizer offset (<= k-m+1 <= 63, src/partitioned_mphf.cpp:
// 183-211), so 32 consecutive entries fit one 32-byte sector:
//   word 0: S[32 s] in bits 0..47, sum of d[32 s .. 32 s + 15] in bits 48..63
//   word 1: low nibbles of d[32 s + 0..15]      word 2: low nibbles of d[32 s + 16..31]
//   word 3: top two bits of d[32 s + 0..31] (2 bits each)
struct DevPrefix {
    const uint64_t* sectors;  // null when some d[i] > 63 or S >= 2^48 (then DevEF is used instead)
    uint64_t n;               // number of S entries
};

struct DevRank {             // rs_bit_vector re-laid-out: unit u = {ones before bit 96*u, 96 bits}
    const uint4* units;      // x = ones before, y/z/w = bits 96u.., 96u+32.., 96u+64..
    uint64_t nbits;          // < 2^32 (PTHash with 64-bit hashes holds <= 2^30 keys, hasher.hpp:27-31)
    uint64_t num_ones;
};

struct DevPhf {              // pthash::single_phf<*, dictionary_dictionary, true>
    uint64_t seed, num_keys, table_size;
    uint64_t dense, sparse;  // skew_bucketer bucket counts (bucketers.hpp:10-22)
    // 96-bit reciprocals ceil(2^96/d) for d < 2^32 (exact a % d for 64-bit a); zero when d >= 2^32
    uint32_t m_table[3], m_dense[3], m_sparse[3];
    uint32_t small_divisors; // 1 when table_size, dense, sparse are all < 2^32
    uint32_t ranks_are_u16;
    const void* ranks;       // one entry per bucket, index into hashed_pilots
    const uint64_t* hashed_pilots;
    DevEF free_slots;        // file layout (pthash ef_sequence<false>)
    const uint32_t* free32;  // the same values decoded (null if one does not fit 32 bits):
                             // free32[i] == free_slots.access(i), single_phf.hpp:61-63
};

struct DevImage {
    uint32_t k, m, w;        // w = k - m + 1
    uint32_t kmer_bits;
    uint64_t mm_seed, nkmers, distinct_minimizers, n_maximal;
    uint64_t right_start, none_sizes_start, none_pos_start;
    uint64_t maximal_block;  // w * n_maximal
    uint64_t collision_base; // EF[none_pos_start] + w*n_maximal (global rank of colliding k-mers)
    DevPhf minimizer_order, fallback;
    DevRank root, left_right, max_none;
    DevEF sp;                // sizes_and_positions (file layout; used when sp_fast.sectors == null)
    DevPrefix sp_fast;       // same values, one 32-byte sector per 32 entries
};

}  // namespace lphb
