#include "lph_image.h"

#include <cstdlib>
#include <cstring>

namespace lphb {

namespace {

// MurmurHash2-64 of one 8-byte word (pthash hasher.hpp:46-110 with len == 8); used at load time
// only, to pre-hash the pilot dictionaries.
inline uint64_t murmur64_host(uint64_t v, uint64_t seed) {
    const uint64_t M = 0xc6a4a7935bd1e995ULL;
    uint64_t h = seed ^ (8 * M);
    uint64_t x = v * M;
    x ^= x >> 47;
    x *= M;
    h = (h ^ x) * M;
    h ^= h >> 47;
    h *= M;
    h ^= h >> 47;
    return h;
}

// A compact_vector still inside the file buffer (its words may be unaligned there).
struct FileCompact {
    uint64_t size = 0, width = 0, mask = 0, nwords = 0;
    const uint8_t* words = nullptr;
    uint64_t word(uint64_t i) const {
        uint64_t v;
        std::memcpy(&v, words + 8 * i, 8);
        return v;
    }
    uint64_t get(uint64_t i) const {
        if (width == 0) return 0;
        uint64_t pos = i * width, wd = pos >> 6, sh = pos & 63;
        uint64_t v = word(wd) >> sh;
        if (sh && wd + 1 < nwords) v |= word(wd + 1) << (64 - sh);
        return v & mask;
    }
};

}  // namespace

struct ImageBuilder::Cursor {
    const uint8_t* p;
    const uint8_t* end;
    template <class T>
    T pod() {
        if (uint64_t(end - p) < sizeof(T)) throw FormatError("truncated .lph image");
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    // vector<T>: returns pointer into the file buffer (possibly unaligned) and its length
    template <class T>
    const uint8_t* vec(uint64_t& n) {
        n = pod<uint64_t>();
        if (n > uint64_t(end - p) / sizeof(T)) throw FormatError("truncated vector in .lph image");
        const uint8_t* q = p;
        p += n * sizeof(T);
        return q;
    }
    FileCompact compact() {
        FileCompact c;
        c.size = pod<uint64_t>();
        c.width = pod<uint64_t>();
        c.mask = pod<uint64_t>();
        c.words = vec<uint64_t>(c.nwords);
        if (c.width > 57) throw FormatError("compact_vector width > 57 unsupported");
        uint64_t expect = c.width ? ((uint64_t(1) << c.width) - 1) : 0;
        if (c.mask != expect) throw FormatError("compact_vector mask/width mismatch");
        if (c.width && (c.size * c.width + 63) / 64 > c.nwords)
            throw FormatError("compact_vector shorter than size*width");
        return c;
    }
};

template <class T>
const T* ImageBuilder::append(const T* src, uint64_t n, uint64_t pad_elems) {
    uint64_t off = (arena_.size() + 255) & ~uint64_t(255);
    arena_.resize(off + (n + pad_elems) * sizeof(T), 0);
    if (n) std::memcpy(arena_.data() + off, src, n * sizeof(T));
    return reinterpret_cast<const T*>(uintptr_t(off));
}

std::vector<uint64_t> ImageBuilder::read_ef(Cursor& c) {
    // pthash ef_sequence / lphash ef_sequence share one layout: {bit_vector high; darray1;
    // compact_vector low} (include/ef_sequence.hpp:107-112, pthash darray.hpp:88-94).
    // value i = ((position of the i-th set high bit) - i) << l | low[i]  (ef_sequence.hpp:77-81)
    uint64_t nbits = c.pod<uint64_t>();
    uint64_t nw;
    const uint8_t* hw = c.vec<uint64_t>(nw);
    if ((nbits + 63) / 64 > nw) throw FormatError("EF high bits shorter than declared");
    uint64_t positions = c.pod<uint64_t>();
    uint64_t nb, ns, no;
    c.vec<int64_t>(nb);   // darray block inventory: select is not needed once decoded
    c.vec<uint16_t>(ns);  // subblock inventory
    c.vec<uint64_t>(no);  // overflow positions
    FileCompact low = c.compact();
    if (low.size != positions) throw FormatError("EF: darray positions != number of values");
    std::vector<uint64_t> vals;
    vals.reserve(positions);
    for (uint64_t wi = 0; wi < nw && vals.size() < positions; ++wi) {
        uint64_t wv;
        std::memcpy(&wv, hw + 8 * wi, 8);
        while (wv && vals.size() < positions) {
            uint64_t pos = wi * 64 + uint64_t(__builtin_ctzll(wv));
            wv &= wv - 1;
            if (pos >= nbits) break;
            uint64_t i = vals.size();
            vals.push_back(((pos - i) << low.width) | low.get(i));
        }
    }
    if (vals.size() != positions) throw FormatError("EF: fewer set bits than values");
    return vals;
}

ImageBuilder::Bits ImageBuilder::read_bits(Cursor& c) {
    // rs_bit_vector = {u64 nbits; vec<u64> bits; vec<u64> block_rank_pairs; vec<u64> select_hints}
    // (include/rs_bit_vector.hpp:91-96)
    Bits out;
    out.nbits = c.pod<uint64_t>();
    uint64_t nw, np, nh;
    const uint8_t* bw = c.vec<uint64_t>(nw);
    c.vec<uint64_t>(np);  // rank directory: ranks are recomputed while the buckets are walked
    c.vec<uint64_t>(nh);  // select hints: never built by the reference (quartet_wtree.cpp:51-53)
    if ((out.nbits + 63) / 64 > nw) throw FormatError("bit vector shorter than declared");
    out.words.resize(nw + 1, 0);
    if (nw) std::memcpy(out.words.data(), bw, nw * 8);
    return out;
}

// One word per bucket id: what mphf::query computes from the bucket alone
// (src/partitioned_mphf.cpp:292-339).  Buckets are walked in id order, so the ranks that
// quartet_wtree::rank_of (src/quartet_wtree.cpp:84-99) obtains from its rank directories are
// plain running counters here.  `sp` = decoded sizes_and_positions prefix sums.
void ImageBuilder::build_buckets(Bits const& root, Bits const& left_right, Bits const& max_none,
                                 std::vector<uint64_t> const& sp, std::vector<uint32_t> const& free_slots) {
    const uint64_t D = img_.distinct_minimizers;
    const uint64_t w = img_.w, maxblock = w * img_.n_maximal;
    const uint64_t rs = img_.right_start, ns = img_.none_sizes_start, np = img_.none_pos_start;
    auto S = [&](uint64_t i) -> uint64_t {
        if (i >= sp.size()) throw FormatError("sizes_and_positions index out of range");
        return sp[i];
    };
    img_.collision_base = S(np) + maxblock;  // partitioned_mphf.cpp:308-311
    // The table is indexed by the PTHash table slot BEFORE the minimal remap: slots >= num_keys
    // repeat the word of the key position free_slots sends them to (single_phf.hpp:61-63), so the
    // query needs no free-slot lookup.
    const uint64_t T = D + free_slots.size();
    std::vector<uint64_t> e(T);
    uint64_t n0 = 0, n1 = 0;              // zeros / ones of root seen so far
    uint64_t r_left = 0, r_right = 0, r_max = 0, r_none = 0;
    uint64_t max_base = 0;
    for (uint64_t b = 0; b < D; ++b) {
        bool msb = root.get(b);
        uint64_t base = 0, kind = 0;
        if (!msb) {
            if (n0 >= left_right.nbits) throw FormatError("wavelet tree: left/right leaf too short");
            bool lsb = left_right.get(n0++);
            if (!lsb) {  // LEFT: g = EF[rank] + w*n_max, l = p
                base = S(r_left++) + maxblock;
                kind = 1;
            } else {     // RIGHT or collision (size 0)
                uint64_t v1 = S(rs + r_right), v2 = S(rs + r_right + 1);
                ++r_right;
                if (v2 == v1) {
                    kind = 0;
                } else {  // g = v1 + w*n_max, l = (k-m) - p
                    base = v1 + maxblock + uint64_t(img_.k - img_.m);
                    kind = 2;
                }
            }
        } else {
            if (n1 >= max_none.nbits) throw FormatError("wavelet tree: max/none leaf too short");
            bool lsb = max_none.get(n1++);
            if (!lsb) {  // MAXIMAL: g = w*rank, l = p
                base = w * r_max++;
                kind = 1;
            } else {     // NONE: g = EF[none_sizes_start+rank] + w*n_max, l = diff(none_pos) - p
                base = S(ns + r_none) + maxblock + (S(np + r_none + 1) - S(np + r_none));
                ++r_none;
                kind = 2;
            }
        }
        if (base >> 62) throw FormatError("bucket base does not fit 62 bits");
        if (base > max_base) max_base = base;
        // bit 63: slope +1 (else -1), bit 62: colliding minimizer
        e[b] = (uint64_t(kind == 1) << 63) | (uint64_t(kind == 0) << 62) | base;
    }
    for (uint64_t i = 0; i < free_slots.size(); ++i) {
        if (free_slots[i] >= D) throw FormatError("single_phf: free slot outside [0, num_keys)");
        e[D + i] = e[free_slots[i]];
    }
    img_.buckets.n = T;
    // 32-bit entries whenever every base fits 30 bits; LPHB_FORCE_WIDE_BUCKETS=1 (test hook) keeps the
    // 64-bit form, which only indexes of >= 2^30 k-mers would otherwise exercise
    const char* fw = getenv("LPHB_FORCE_WIDE_BUCKETS");
    const bool force_wide = fw && fw[0] && fw[0] != '0';
    if (max_base < (1ull << 30) && !force_wide) {
        std::vector<uint32_t> e32(T);
        for (uint64_t b = 0; b < T; ++b)
            e32[b] = uint32_t(e[b] >> 62) << 30 | uint32_t(e[b] & 0x3FFFFFFFull);  // same two flag bits on top
        img_.buckets.wide = 0;
        img_.buckets.entries = append(e32.data(), T, 8);
    } else {
        img_.buckets.wide = 1;
        img_.buckets.entries = append(e.data(), T, 4);
    }
}

void reciprocal64(uint64_t d, uint32_t out[2]) {
    // M = floor(2^64 / d); d == 1 uses 2^64 - 1 (mod_small still returns 0)
    unsigned __int128 M = d > 1 ? ((((unsigned __int128)1) << 64) / d) : ((((unsigned __int128)1) << 64) - 1);
    out[0] = uint32_t(M);
    out[1] = uint32_t(M >> 32);
}

void ImageBuilder::read_phf(Cursor& c, DevPhf& out) {
    out.seed = c.pod<uint64_t>();
    out.num_keys = c.pod<uint64_t>();
    out.table_size = c.pod<uint64_t>();
    c.pod<unsigned __int128>();  // M: a % d is computed exactly on the device without it
    out.dense = c.pod<uint64_t>();
    out.sparse = c.pod<uint64_t>();
    c.pod<unsigned __int128>();
    c.pod<unsigned __int128>();
    FileCompact fr = c.compact(), fd = c.compact(), br = c.compact(), bd = c.compact();
    uint64_t nbuckets = fr.size + br.size, ndict = fd.size + bd.size;
    if (nbuckets != out.dense + out.sparse) throw FormatError("pilot count != bucket count");
    if (out.table_size == 0 || out.table_size < out.num_keys)
        throw FormatError("single_phf: bad table size");
    if (out.dense == 0 || out.sparse == 0) throw FormatError("single_phf: empty bucket class");
    // one word per bucket: default_hash64(pilot of the bucket, seed) (single_phf.hpp:57-58), the
    // dictionary indirection of dual<dictionary, dictionary> resolved here
    std::vector<uint64_t> hp(ndict);
    for (uint64_t i = 0; i < fd.size; ++i) hp[i] = murmur64_host(fd.get(i), out.seed);
    for (uint64_t i = 0; i < bd.size; ++i) hp[fd.size + i] = murmur64_host(bd.get(i), out.seed);
    std::vector<uint64_t> per_bucket(nbuckets);
    for (uint64_t b = 0; b < nbuckets; ++b) {
        uint64_t r;
        if (b < fr.size) {
            r = fr.get(b);
            if (r >= fd.size) throw FormatError("pilot rank outside dictionary");
        } else {
            r = br.get(b - fr.size);
            if (r >= bd.size) throw FormatError("pilot rank outside dictionary");
            r += fd.size;
        }
        per_bucket[b] = hp[r];
    }
    out.pilot_hash = append(per_bucket.data(), nbuckets, 2);
    // 64-bit PTHash hashes cap num_keys at 2^30 (hasher.hpp:27-31), so table_size = num_keys/alpha
    // stays below 2^31, which the device-side exact modulo relies on
    if (out.table_size >= (1ull << 31) || out.dense >= (1ull << 31) || out.sparse >= (1ull << 31))
        throw FormatError("single_phf: table larger than 2^31 (impossible with 64-bit PTHash hashes)");
    reciprocal64(out.table_size, out.m_table);
    reciprocal64(out.dense, out.m_dense);
    reciprocal64(out.sparse, out.m_sparse);
    // minimal remap as a plain array: one load instead of an Elias-Fano select
    std::vector<uint64_t> free_vals = read_ef(c);
    if (free_vals.size() != out.table_size - out.num_keys)
        throw FormatError("single_phf: free-slot count mismatch");
    std::vector<uint32_t> f32(free_vals.size());
    for (size_t i = 0; i < free_vals.size(); ++i) {
        if (free_vals[i] >= out.table_size) throw FormatError("single_phf: free slot outside the table");
        f32[i] = uint32_t(free_vals[i]);
    }
    out.free32 = append(f32.data(), f32.size(), 4);
    last_free_ = std::move(f32);
}

void ImageBuilder::parse(const uint8_t* data, uint64_t n, int kmer_bits) {
    if (kmer_bits != 64 && kmer_bits != 128) throw FormatError("kmer_bits must be 64 or 128");
    arena_.clear();
    img_ = DevImage{};
    Cursor c{data, data + n};
    img_.k = c.pod<uint8_t>();
    img_.m = c.pod<uint8_t>();
    img_.kmer_bits = uint32_t(kmer_bits);
    img_.mm_seed = c.pod<uint64_t>();
    img_.nkmers = c.pod<uint64_t>();
    img_.distinct_minimizers = c.pod<uint64_t>();
    img_.n_maximal = c.pod<uint64_t>();
    img_.right_start = c.pod<uint64_t>();
    img_.none_sizes_start = c.pod<uint64_t>();
    img_.none_pos_start = c.pod<uint64_t>();
    if (img_.m == 0 || img_.m > 31 || img_.k < img_.m || img_.k > uint32_t(kmer_bits / 2 - 1))
        throw FormatError("k/m out of range for this kmer_t");
    img_.w = img_.k - img_.m + 1;
    read_phf(c, img_.minimizer_order);
    std::vector<uint32_t> mo_free = std::move(last_free_);
    Bits root = read_bits(c), left_right = read_bits(c), max_none = read_bits(c);
    std::vector<uint64_t> sp = read_ef(c);
    read_phf(c, img_.fallback);
    if (c.p != c.end) throw FormatError("trailing bytes after the .lph image");
    if (img_.minimizer_order.num_keys != img_.distinct_minimizers ||
        root.nbits != img_.distinct_minimizers ||
        left_right.nbits + max_none.nbits != img_.distinct_minimizers)
        throw FormatError("inconsistent minimizer counts");
    if (!(img_.right_start <= img_.none_sizes_start && img_.none_sizes_start <= img_.none_pos_start &&
          img_.none_pos_start < sp.size()))
        throw FormatError("inconsistent sizes_and_positions partition");
    build_buckets(root, left_right, max_none, sp, mo_free);
    fallback_keys_ = img_.fallback.num_keys;
    file_bytes_ = n;
    arena_.resize((arena_.size() + 255) & ~uint64_t(255), 0);
}

DevImage ImageBuilder::rebased(const void* device_base) const {
    DevImage d = img_;
    auto* base = static_cast<const uint8_t*>(device_base);
    for (DevPhf* p : {&d.minimizer_order, &d.fallback}) {
        p->free32 = reinterpret_cast<const uint32_t*>(base + uintptr_t(p->free32));
        p->pilot_hash = reinterpret_cast<const uint64_t*>(base + uintptr_t(p->pilot_hash));
    }
    d.buckets.entries = base + uintptr_t(d.buckets.entries);
    return d;
}

}  // namespace lphb
