// lphash_b200.hpp — C++ host-side mirror of the reference's interface for the hot path, on top of
// the C ABI (include/lphash_b200.h -> lphash_b200/liblphash_b200.so).  Header-only; C++17.
//
// The reference is compiled C++ whose seam for this path is a member function and two free
// function templates, so the drop-in keeps their names, argument meaning and error behaviour:
//
//   reference (jermp/lphash)                                   here (namespace lphash_b200)
//   ---------------------------------------------------------  -------------------------------------
//   lphash::mphf hf; essentials::load(hf, file)                mphf hf; hf.load(file)      (query.cpp:36-38)
//   hf(contig, length, streaming=true) -> vector<uint64_t>     same   (partitioned_mphf.hpp:21-26, 73-184)
//   hf(std::string const&, streaming=true)                     same   (partitioned_mphf.hpp:199-202)
//   hf.get_kmer_count(), hf.get_minimizer_L0()                 same   (partitioned_mphf.hpp:17-18)
//   minimizer::from_string<H>(contig, n, k, m, seed,           minimizer::from_string(...) same argument
//       canonical, mm_count, accumulator)                      list; accumulator = anything with
//                                                              push_back(mm_record_t)  (minimizer.hpp:11-170)
//   minimizer::get_colliding_kmers<H>(contig, n, k, m, seed,   minimizer::get_colliding_kmers(...) same
//       canonical, itr, stop, mm_count, accumulator)           argument list           (minimizer.hpp:172-319)
//
// Errors surface as std::runtime_error, like the reference (partitioned_mphf.cpp:67, :335).  There
// is no CPU path behind this header: without the CUDA library / a GPU every call throws.
//
// kmer_t follows the reference's compile-time switch (include/compile_constants.tpd): define
// LPHASH_B200_KMER_BITS to 64 for a `typedef uint64_t kmer_t;` build; the default is 128
// (`__uint128_t`, the reference's default).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "lphash_b200.h"

#ifndef LPHASH_B200_KMER_BITS
#define LPHASH_B200_KMER_BITS 128
#endif

namespace lphash_b200 {

#if LPHASH_B200_KMER_BITS == 64
typedef uint64_t kmer_t;
#else
typedef __uint128_t kmer_t;
#endif
constexpr int kmer_bits = LPHASH_B200_KMER_BITS;

// include/constants.hpp:26-33
#pragma pack(push, 2)
struct mm_record_t {
    uint64_t itself;
    uint64_t id;
    uint8_t p1;
    uint8_t size;
};
#pragma pack(pop)
static_assert(sizeof(mm_record_t) == 18, "mm_record_t is the reference's packed 18-byte record");

// include/constants.hpp:35-41
#pragma pack(push, 2)
struct mm_triplet_t {
    uint64_t itself;
    uint8_t p1;
    uint8_t size;
};
#pragma pack(pop)
static_assert(sizeof(mm_triplet_t) == 10, "mm_triplet_t is the reference's packed 10-byte triplet");

namespace detail {
[[noreturn]] inline void raise(const char* what, int rc) {
    throw std::runtime_error(std::string(what) + ": " + lphb_last_error() + " (code " +
                             std::to_string(rc) + ")");
}
}  // namespace detail

class mphf {
public:
    mphf() = default;
    mphf(mphf const&) = delete;
    mphf& operator=(mphf const&) = delete;
    mphf(mphf&& o) noexcept : h_(o.h_), info_(o.info_) { o.h_ = nullptr; }
    mphf& operator=(mphf&& o) noexcept {
        if (this != &o) {
            reset();
            h_ = o.h_;
            info_ = o.info_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ~mphf() { reset(); }

    // essentials::load(hf, filename) (src/query.cpp:37).  `device` = CUDA device holding the image.
    void load(const char* filename, int device = 0) {
        reset();
        int rc = lphb_mphf_load_file(filename, kmer_bits, device, &h_);
        if (rc != LPHB_OK) detail::raise("lphash_b200::mphf::load", rc);
        lphb_mphf_info(h_, &info_);
    }
    // the same from an in-memory essentials image (what essentials::save writes)
    void load(const void* image, uint64_t nbytes, int device = 0) {
        reset();
        int rc = lphb_mphf_load_memory(image, nbytes, kmer_bits, device, &h_);
        if (rc != LPHB_OK) detail::raise("lphash_b200::mphf::load", rc);
        lphb_mphf_info(h_, &info_);
    }

    // the unpartitioned variant (lphash::mphf_alt, build-u / query-u): same handle, same queries
    void load_alt(const char* filename, int device = 0) {
        reset();
        int rc = lphb_mphf_alt_load_file(filename, kmer_bits, device, &h_);
        if (rc != LPHB_OK) detail::raise("lphash_b200::mphf::load_alt", rc);
        lphb_mphf_info(h_, &info_);
    }
    void load_alt(const void* image, uint64_t nbytes, int device = 0) {
        reset();
        int rc = lphb_mphf_alt_load_memory(image, nbytes, kmer_bits, device, &h_);
        if (rc != LPHB_OK) detail::raise("lphash_b200::mphf::load_alt", rc);
        lphb_mphf_info(h_, &info_);
    }

    uint64_t get_minimizer_L0() const noexcept { return info_.distinct_minimizers; }
    uint64_t get_kmer_count() const noexcept { return info_.nkmers; }
    uint32_t get_k() const noexcept { return info_.k; }
    uint32_t get_m() const noexcept { return info_.m; }
    lphb_info const& info() const noexcept { return info_; }

    // hf(contig, length, streaming).  streaming = true: the reference's streaming branch, non-ACGT quirk
    // included (partitioned_mphf.hpp:73-184); false: its non-streaming branch, where a non-ACGT byte
    // counts as 'A' (partitioned_mphf.hpp:185-195, mphf_utils.hpp:108).  Same codes on ACGT-only input.
    std::vector<uint64_t> operator()(const char* contig, std::size_t length, bool streaming = true) const {
        const uint64_t offsets[2] = {0, uint64_t(length)};
        std::vector<uint64_t> codes, code_offsets;
        query_batch(contig, offsets, 1, codes, code_offsets, streaming);
        return codes;
    }
    std::vector<uint64_t> operator()(std::string const& contig, bool streaming = true) const {
        return (*this)(contig.c_str(), contig.length(), streaming);
    }

    // Many contigs per call (what a throughput-minded driver uses instead of one call per kseq
    // record): bases = contigs concatenated without separators, offsets = n_contigs + 1 entries.
    // codes = all contigs' codes in contig order; contig c owns
    // [code_offsets[c], code_offsets[c+1]).
    void query_batch(const char* bases, const uint64_t* offsets, uint64_t n_contigs,
                     std::vector<uint64_t>& codes, std::vector<uint64_t>& code_offsets,
                     bool streaming = true) const {
        if (!h_) throw std::runtime_error("lphash_b200::mphf: no index loaded");
        // a contig with non-ACGT bytes can emit up to L-m+1 codes (reference quirk)
        uint64_t cap = 0;
        for (uint64_t c = 0; c < n_contigs; ++c) {
            uint64_t len = offsets[c + 1] - offsets[c];
            cap += len >= info_.m ? len - info_.m + 1 : 0;
        }
        codes.resize(cap);
        code_offsets.resize(n_contigs + 1);
        uint64_t n = 0;
        int rc = streaming ? lphb_query_stream(h_, bases, offsets, n_contigs, codes.data(), cap, code_offsets.data(), &n)
                           : lphb_query_nonstreaming(h_, bases, offsets, n_contigs, codes.data(), cap,
                                                     code_offsets.data(), &n);
        if (rc != LPHB_OK) detail::raise("lphash_b200::mphf::operator()", rc);
        codes.resize(n);
    }

    lphb_mphf* handle() const noexcept { return h_; }

private:
    void reset() noexcept {
        if (h_) lphb_mphf_free(h_);
        h_ = nullptr;
    }
    lphb_mphf* h_ = nullptr;
    lphb_info info_{};
};

// Serializes any object that speaks the essentials visitor protocol (`template <class V> void visit(V&)`)
// into memory, byte for byte what essentials::save writes to a file (essentials.hpp:326-366: PODs raw,
// std::vector<T> = size_t count + elements): the `.lph` image of a reference lphash::mphf without a
// round trip through the file system.  mphf::load(image, nbytes) takes the result.
struct memory_saver {
    std::vector<unsigned char> bytes;
    template <typename T>
    void visit(T& val) {
        if constexpr (std::is_pod<T>::value) {
            const unsigned char* p = reinterpret_cast<const unsigned char*>(&val);
            bytes.insert(bytes.end(), p, p + sizeof(T));
        } else {
            val.visit(*this);
        }
    }
    template <typename T, typename Allocator>
    void visit(std::vector<T, Allocator>& vec) {
        size_t n = vec.size();
        visit(n);
        if constexpr (std::is_pod<T>::value) {
            const unsigned char* p = reinterpret_cast<const unsigned char*>(vec.data());
            bytes.insert(bytes.end(), p, p + n * sizeof(T));
        } else {
            for (auto& v : vec) visit(v);
        }
    }
};

// The inverse: essentials' loader protocol (essentials.hpp:280-324) over a byte image in memory.
struct memory_loader {
    const unsigned char* p;
    const unsigned char* end;
    memory_loader(const void* image, size_t nbytes)
        : p(static_cast<const unsigned char*>(image)), end(static_cast<const unsigned char*>(image) + nbytes) {}
    template <typename T>
    void visit(T& val) {
        if constexpr (std::is_pod<T>::value) {
            take(&val, sizeof(T));
        } else {
            val.visit(*this);
        }
    }
    template <typename T, typename Allocator>
    void visit(std::vector<T, Allocator>& vec) {
        size_t n = 0;
        visit(n);
        if constexpr (std::is_pod<T>::value) {
            if (n > size_t(end - p) / sizeof(T)) throw std::runtime_error("lphash_b200: truncated image");
            vec.resize(n);
            take(vec.data(), n * sizeof(T));
        } else {
            vec.resize(n);
            for (auto& v : vec) visit(v);
        }
    }

private:
    void take(void* dst, size_t n) {
        if (size_t(end - p) < n) throw std::runtime_error("lphash_b200: truncated image");
        if (n) std::memcpy(dst, p, n);
        p += n;
    }
};

namespace minimizer {

// minimizer::from_string (include/minimizer.hpp:11-170): appends one mm_record_t per super-k-mer
// of `contig` to `accumulator` (anything with push_back(mm_record_t), e.g. the reference's
// external_memory_vector<mm_record_t>), advances the running m-mer ordinal `mm_count`, returns the
// number of k-mers.  `canonical_m_mers` must be false (the reference never passes true:
// src/partitioned_mphf.cpp:34).
template <class Accumulator>
[[nodiscard]] uint64_t from_string(char const* contig, std::size_t contig_size, uint32_t k,
                                   uint32_t m, uint64_t seed, bool canonical_m_mers,
                                   uint64_t& mm_count, Accumulator& accumulator, int device = 0) {
    if (canonical_m_mers) throw std::runtime_error("lphash_b200: canonical m-mers are not supported");
    const uint64_t offsets[2] = {0, uint64_t(contig_size)};
    const uint64_t cap = contig_size >= k ? contig_size - k + 1 : 0;
    std::vector<mm_record_t> rec(cap ? cap : 1);
    uint64_t n_rec = 0, n_kmers = 0;
    int rc = lphb_scan_superkmers(device, k, m, seed, contig, offsets, 1, &mm_count, rec.data(),
                                  cap, &n_rec, &n_kmers);
    if (rc != LPHB_OK) detail::raise("lphash_b200::minimizer::from_string", rc);
    for (uint64_t i = 0; i < n_rec; ++i) accumulator.push_back(rec[i]);
    return n_kmers;
}

// Sort by minimizer + minimizer::classify (src/minimizer.cpp:5-50): unique minimizers in ascending
// order as triplets ({itself, 0, 0} for minimizers seen several times) and the ascending ids of the
// occurrences of the latter.  `records` may be in any order (scan order is fine).
inline void classify(std::vector<mm_record_t> const& records, std::vector<mm_triplet_t>& unique_minimizers,
                     std::vector<uint64_t>& colliding_minimizer_ids, int device = 0) {
    const uint64_t n = records.size();
    unique_minimizers.resize(n ? n : 1);
    colliding_minimizer_ids.resize(n ? n : 1);
    uint64_t nt = 0, ni = 0;
    int rc = lphb_classify(device, records.data(), n, unique_minimizers.data(), n, &nt,
                           colliding_minimizer_ids.data(), n, &ni);
    if (rc != LPHB_OK) detail::raise("lphash_b200::minimizer::classify", rc);
    unique_minimizers.resize(nt);
    colliding_minimizer_ids.resize(ni);
}

// Frees the device workspace the two build-side calls keep between calls (call after the loop over
// from_string / get_colliding_kmers, i.e. after src/partitioned_mphf.cpp:77 and :129).
inline void release_workspace(int device = 0) {
    int rc = lphb_scan_release(device);
    if (rc != LPHB_OK) detail::raise("lphash_b200::minimizer::release_workspace", rc);
}

// Batch form: the whole input (or a large slab of it) in one call; records land in `records` in
// scan order.  Returns the number of k-mers.
inline uint64_t from_batch(const char* bases, const uint64_t* offsets, uint64_t n_contigs, uint32_t k,
                           uint32_t m, uint64_t seed, uint64_t& mm_count,
                           std::vector<mm_record_t>& records, int device = 0) {
    uint64_t cap = 0;
    for (uint64_t c = 0; c < n_contigs; ++c) {
        uint64_t len = offsets[c + 1] - offsets[c];
        cap += len >= k ? len - k + 1 : 0;
    }
    records.resize(cap ? cap : 1);
    uint64_t n_rec = 0, n_kmers = 0;
    int rc = lphb_scan_superkmers(device, k, m, seed, bases, offsets, n_contigs, &mm_count,
                                  records.data(), cap, &n_rec, &n_kmers);
    if (rc != LPHB_OK) detail::raise("lphash_b200::minimizer::from_batch", rc);
    records.resize(n_rec);
    return n_kmers;
}

// minimizer::get_colliding_kmers (include/minimizer.hpp:172-319): appends to `accumulator` every
// k-mer of the super-k-mers of `contig` whose minimizer-occurrence id appears in the ascending id
// stream [itr, stop); advances `itr` past the ids consumed and `mm_count` past the contig's m-mers.
template <class IdIterator, class Accumulator>
void get_colliding_kmers(char const* contig, std::size_t contig_size, uint32_t k, uint32_t m,
                         uint64_t seed, bool canonical_m_mers, IdIterator& itr, IdIterator& stop,
                         uint64_t& mm_count, Accumulator& accumulator, int device = 0) {
    if (canonical_m_mers) throw std::runtime_error("lphash_b200: canonical m-mers are not supported");
    const uint64_t n_mmers = contig_size >= m ? contig_size - m + 1 : 0;
    const uint64_t id_end = mm_count + n_mmers;
    // ids that can belong to this contig: [mm_count, id_end)
    std::vector<uint64_t> ids;
    IdIterator probe = itr;
    while (probe != stop && *probe < id_end) {
        ids.push_back(*probe);
        ++probe;
    }
    const uint64_t offsets[2] = {0, uint64_t(contig_size)};
    const uint64_t cap = contig_size >= k ? contig_size - k + 1 : 0;
    std::vector<kmer_t> out(cap ? cap : 1);
    uint64_t n = 0;
    int rc = lphb_colliding_kmers(device, k, m, seed, contig, offsets, 1, &mm_count, ids.data(),
                                  ids.size(), kmer_bits, out.data(), cap, &n);
    if (rc != LPHB_OK) detail::raise("lphash_b200::minimizer::get_colliding_kmers", rc);
    for (uint64_t i = 0; i < n; ++i) accumulator.push_back(out[i]);
    itr = probe;
}

}  // namespace minimizer

}  // namespace lphash_b200
