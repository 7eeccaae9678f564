// Flat device image of one serialized lphash::mphf — the HBM-resident form of the structures the
// query probes.  POD only: passed to kernels by value (__grid_constant__).
//
// The image is NOT the serialized layout.  It holds the same function re-laid-out for the GPU
// (results are bit-identical; lph_image.cpp does the transformations once, at load time):
//   * pilots: the reference's dual<dictionary,dictionary> (pthash encoders.hpp:167-176, 268-277)
//     is two (ranks, dict) compact-vector pairs; here every bucket directly holds
//     default_hash64(pilot, seed) (single_phf.hpp:58): ONE 8-byte gather per probe instead of two
//     dependent ones, and no second murmur.
//   * free slots: pthash's Elias-Fano `free_slots` (single_phf.hpp:61-63) decoded into a plain
//     u32 array (table_size < 2^32 because 64-bit PTHash hashes cap num_keys at 2^30,
//     hasher.hpp:27-31).
//   * bucket table: quartet_wtree::rank_of + the sizes_and_positions Elias-Fano lookups of
//     mphf::query (src/partitioned_mphf.cpp:292-339; src/quartet_wtree.cpp:84-99;
//     include/ef_sequence.hpp:77-99) depend only on the bucket id (the minimizer's MPHF value),
//     so they are evaluated for every bucket once and stored as one word per bucket:
//         hval(k-mer) = base + slope * p      p = offset of the minimizer inside the k-mer
//     entry = flags | base; top bit: slope +1 (LEFT, MAXIMAL) else -1 (RIGHT, NONE); next bit:
//     colliding minimizer (base unused; every k-mer goes through fallback_kmer_order).
//     32-bit entries when every base < 2^30, else 64-bit.  One gather per super-k-mer instead of
//     two rank lookups and up to three Elias-Fano accesses.  The table has table_size entries,
//     indexed by the PTHash slot BEFORE the minimal remap: a slot >= num_keys repeats the word of
//     the position free_slots maps it to, so no free-slot lookup is needed either.
#pragma once
#include <stdint.h>

namespace lphb {

struct DevPhf {              // pthash::single_phf<*, dictionary_dictionary, true>
    uint64_t seed, num_keys, table_size;
    uint64_t dense, sparse;  // skew_bucketer bucket counts (bucketers.hpp:10-22); all < 2^31
    // 64-bit reciprocals floor(2^64/d) as two 32-bit limbs (exact a % d for 64-bit a and d < 2^31,
    // device_mphf.cuh: mod_small)
    uint32_t m_table[2], m_dense[2], m_sparse[2];
    const uint64_t* pilot_hash;  // per bucket: default_hash64(pilot, seed)
    const uint32_t* free32;  // free32[i] == free_slots.access(i)
};

struct DevBuckets {          // per PTHash table slot: flags | base (see above)
    const void* entries;
    uint32_t wide;           // 0: uint32_t entries (flags in bits 30-31), 1: uint64_t (bits 62-63)
    uint64_t n;
};

struct DevImage {
    uint32_t k, m, w;        // w = k - m + 1
    uint32_t kmer_bits;
    uint64_t mm_seed, nkmers, distinct_minimizers, n_maximal;
    uint64_t right_start, none_sizes_start, none_pos_start;
    uint64_t collision_base; // EF[none_pos_start] + w*n_maximal (global rank of colliding k-mers)
    DevPhf minimizer_order, fallback;
    DevBuckets buckets;
};

}  // namespace lphb
