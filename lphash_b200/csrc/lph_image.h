// Host side of the device image: parses the serialized lphash::mphf (`.lph`) and lays the arrays
// out in one arena that is uploaded with a single copy.
//
// File format (little-endian, essentials visitor: PODs raw, vector<T> = u64 n + n*sizeof(T)):
//   u8 k, u8 m, u64 mm_seed, nkmers, distinct_minimizers, n_maximal, right_coll_sizes_start,
//   none_sizes_start, none_pos_start            ref include/partitioned_mphf.hpp:204-211
//   single_phf minimizer_order                  ref pthash/include/single_phf.hpp:88-97
//   quartet_wtree {root, left_right, max_none}  ref include/quartet_wtree.hpp:43-48,
//                                                   include/rs_bit_vector.hpp:91-96
//   ef_sequence sizes_and_positions             ref include/ef_sequence.hpp:107-112
//   single_phf fallback_kmer_order
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "device_image.h"

namespace lphb {

struct FormatError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- what a serialized index consists of, as references into the file buffer (nothing decoded) ----------------
// The device-side loader (image_decode.cu) uploads these pieces and evaluates compact_vector::access, the Elias-Fano
// select / access, rs_bit_vector::rank and quartet_wtree::rank_of on the GPU; ImageBuilder::plan only walks the
// structure and checks what can be checked from sizes alone.
struct VecRef {      // a vector of 8-byte words inside the file (possibly unaligned there)
    const uint8_t* p = nullptr;
    uint64_t n = 0;
};
struct CompactRef {  // pthash::compact_vector
    uint64_t size = 0, width = 0;
    VecRef words;
};
struct EfRef {       // {bit_vector high; darray1 (skipped); compact_vector low}
    uint64_t nbits = 0, positions = 0;
    VecRef high;
    CompactRef low;
};
struct BitsRef {     // rs_bit_vector (rank directory skipped)
    uint64_t nbits = 0;
    VecRef words;
};
struct PhfRef {      // pthash::single_phf<*, dictionary_dictionary, true>
    CompactRef front_ranks, front_dict, back_ranks, back_dict;
    EfRef free_slots;
    uint64_t n_buckets = 0, n_free = 0;
};
struct ImagePlan {
    bool alt = false;
    DevImage img{};           // header fields, reciprocals; array pointers = byte offsets inside the arena
    uint64_t arena_bytes = 0;
    PhfRef minimizer_order, fallback;
    BitsRef root, left_right, max_none;  // partitioned
    EfRef sizes_and_positions;           // partitioned
    EfRef positions, sizes;              // mphf_alt
    uint64_t fallback_keys = 0, file_bytes = 0;
    uint64_t sections[5] = {0, 0, 0, 0, 0};  // as ImageBuilder::sections()
};

// Arena under construction: every array is appended 256-byte aligned; device pointers in the
// DevImage are stored as byte offsets first and rebased once the arena is on the device.
class ImageBuilder {
public:
    // Parses `data[0..n)`; throws FormatError.  After this, image() holds offsets-as-pointers.
    void parse(const uint8_t* data, uint64_t n, int kmer_bits);
    // The same for a serialized lphash::mphf_alt (the unpartitioned variant, build-u / query-u):
    //   u8 k, u8 m, u64 mm_seed, nkmers, distinct_minimizers, num_kmers_in_main_index
    //   single_phf minimizer_order, ef_sequence positions, ef_sequence sizes, single_phf fallback_kmer_order
    //   (ref include/unpartitioned_mphf.hpp:198-210).  Its query (ref src/unpartitioned_mphf.cpp:191-206:
    //   hval = sizes[i] + positions.diff(i) - p, or num_kmers_in_main_index + fallback(kmer) when the size is 0)
    //   fits the same per-bucket word, so the device image and every kernel are shared with the partitioned form.
    void parse_alt(const uint8_t* data, uint64_t n, int kmer_bits);
    // Structure walk only (see ImagePlan): the same checks on sizes and ranges as parse / parse_alt, the same arena
    // layout, no decoding.  The references point into `data`, which must outlive the plan.
    static ImagePlan plan(const uint8_t* data, uint64_t n, int kmer_bits, bool alt);
    // the same for a serialized pthash::single_phf on its own (build-p Part 3): plan.minimizer_order only
    static ImagePlan plan_phf(const uint8_t* data, uint64_t n);
    // A serialized pthash::single_phf on its own (what build-p Part 3 evaluates, ref src/partitioned_mphf.cpp:
    // 96-100): afterwards the image holds `minimizer_order` only, free slots included.
    void parse_phf(const uint8_t* data, uint64_t n);
    // Byte offsets inside the image parsed last: start of minimizer_order, of the wavelet tree (alt: positions),
    // of sizes_and_positions (alt: sizes), of fallback_kmer_order, end of the image.
    const uint64_t* sections() const { return sections_; }
    std::vector<uint8_t> const& arena() const { return arena_; }
    // Returns the image with every pointer rebased onto `device_base`.
    DevImage rebased(const void* device_base) const;
    uint64_t fallback_keys() const { return fallback_keys_; }
    uint64_t file_bytes() const { return file_bytes_; }

private:
    struct Cursor;
    template <class T>
    const T* append(const T* src, uint64_t n, uint64_t pad_elems = 0);
    // every value of a serialized Elias-Fano sequence (pthash layout), decoded
    std::vector<uint64_t> read_ef(Cursor& c);
    // the bits of a serialized rs_bit_vector (rank directory skipped)
    struct Bits {
        uint64_t nbits = 0;
        std::vector<uint64_t> words;
        bool get(uint64_t i) const { return (words[i >> 6] >> (i & 63)) & 1; }
    };
    Bits read_bits(Cursor& c);
    void read_phf(Cursor& c, DevPhf& out);
    static void walk_compact(Cursor& c, CompactRef& r);
    static void walk_ef(Cursor& c, EfRef& r);
    static void walk_phf(Cursor& c, PhfRef& r, DevPhf& out);
    // the bucket table under construction, inside the arena (flag bits on top of the base: device_image.h)
    struct BucketWriter {
        uint8_t* p = nullptr;
        bool wide = false;
        void set(uint64_t i, bool slope_up, bool colliding, uint64_t base) {
            if (wide) reinterpret_cast<uint64_t*>(p)[i] = (uint64_t(slope_up) << 63) | (uint64_t(colliding) << 62) | base;
            else reinterpret_cast<uint32_t*>(p)[i] = (uint32_t(slope_up) << 31) | (uint32_t(colliding) << 30) | uint32_t(base & 0x3FFFFFFFu);
        }
        void copy(uint64_t to, uint64_t from) {
            if (wide) reinterpret_cast<uint64_t*>(p)[to] = reinterpret_cast<uint64_t*>(p)[from];
            else reinterpret_cast<uint32_t*>(p)[to] = reinterpret_cast<uint32_t*>(p)[from];
        }
    };
    BucketWriter begin_buckets(uint64_t T);
    void finish_buckets(BucketWriter& e, uint64_t D, uint64_t max_base, std::vector<uint32_t> const& free_slots);
    void build_buckets(Bits const& root, Bits const& left_right, Bits const& max_none,
                       std::vector<uint64_t> const& sp, std::vector<uint32_t> const& free_slots);

    std::vector<uint8_t> arena_;
    DevImage img_{};
    uint64_t fallback_keys_ = 0, file_bytes_ = 0;
    uint64_t sections_[5] = {0, 0, 0, 0, 0};
    std::vector<uint32_t> last_free_;  // free slots of the PHF parsed last
};

// ceil(2^96 / d) as three 32-bit limbs (d >= 1, d < 2^32); limbs all zero for d == 1.
void reciprocal96(uint64_t d, uint32_t out[3]);

}  // namespace lphb
