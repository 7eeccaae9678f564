#include "lph_image.h"

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

void ImageBuilder::read_compact(Cursor& c, DevCompact& out) {
    FileCompact fc = c.compact();
    out.size = fc.size;
    out.width = uint32_t(fc.width);
    out.mask = fc.mask;
    // two pad words: the device reads words [i, i+1] unconditionally
    uint64_t off = (arena_.size() + 255) & ~uint64_t(255);
    arena_.resize(off + (fc.nwords + 2) * 8, 0);
    if (fc.nwords) std::memcpy(arena_.data() + off, fc.words, fc.nwords * 8);
    out.bits = reinterpret_cast<const uint64_t*>(uintptr_t(off));
}

void ImageBuilder::read_ef(Cursor& c, DevEF& out, DevPrefix* fast, std::vector<uint64_t>* decoded) {
    uint64_t nbits = c.pod<uint64_t>();
    uint64_t nw;
    const uint8_t* hw = c.vec<uint64_t>(nw);
    if ((nbits + 63) / 64 > nw) throw FormatError("EF high bits shorter than declared");
    {
        uint64_t off = (arena_.size() + 255) & ~uint64_t(255);
        arena_.resize(off + (nw + 2) * 8, 0);
        if (nw) std::memcpy(arena_.data() + off, hw, nw * 8);
        out.high = reinterpret_cast<const uint64_t*>(uintptr_t(off));
    }
    uint64_t positions = c.pod<uint64_t>();
    uint64_t nb, ns, no;
    const uint8_t* bi = c.vec<int64_t>(nb);
    const uint8_t* si = c.vec<uint16_t>(ns);
    const uint8_t* ov = c.vec<uint64_t>(no);
    if (nb < (positions + 1023) / 1024 || ns < (positions + 31) / 32)
        throw FormatError("darray inventories shorter than declared");
    out.block_inv = append(reinterpret_cast<const int64_t*>(bi), nb, 1);
    out.sub_inv = append(reinterpret_cast<const uint16_t*>(si), ns, 4);
    out.overflow = append(reinterpret_cast<const uint64_t*>(ov), no, 1);
    FileCompact low_view;
    {
        Cursor peek = c;  // the low-bits vector is parsed twice: once as a view, once into the arena
        low_view = peek.compact();
    }
    read_compact(c, out.low);
    out.n = out.low.size;
    if (out.n != positions) throw FormatError("EF: darray positions != number of values");
    if (!fast && !decoded) return;
    // Decode every value once (value i = ((position of the i-th one) - i) << l | low[i],
    // include/ef_sequence.hpp:77-81) and re-encode as prefix sectors (device_image.h).
    if (fast) {
        fast->sectors = nullptr;
        fast->n = out.n;
    }
    std::vector<uint64_t> vals;
    vals.reserve(out.n);
    for (uint64_t wi = 0; wi < nw && vals.size() < out.n; ++wi) {
        uint64_t wv;
        std::memcpy(&wv, hw + 8 * wi, 8);
        while (wv && vals.size() < out.n) {
            uint64_t pos = wi * 64 + uint64_t(__builtin_ctzll(wv));
            wv &= wv - 1;
            if (pos >= nbits) break;
            uint64_t i = vals.size();
            vals.push_back(((pos - i) << low_view.width) | low_view.get(i));
        }
    }
    if (vals.size() != out.n) throw FormatError("EF: fewer set bits than values");
    if (decoded) *decoded = vals;
    if (!fast) return;
    uint64_t nsec = out.n / 32 + 1;
    std::vector<uint64_t> sec(nsec * 4, 0);
    for (uint64_t s0 = 0; s0 < nsec; ++s0) {
        uint64_t first = s0 * 32;
        uint64_t base = first < out.n ? vals[first] : (out.n ? vals[out.n - 1] : 0);
        if (base >> 48) return;  // does not fit: keep the EF path
        uint64_t half = 0, nib[2] = {0, 0}, top = 0;
        for (uint64_t j = 0; j < 32; ++j) {
            uint64_t i = first + j;
            uint64_t d = 0;
            if (i + 1 < out.n) {
                if (vals[i + 1] < vals[i]) throw FormatError("EF: values not monotone");
                d = vals[i + 1] - vals[i];
            }
            if (d > 63) return;  // not a sizes/positions array after all: keep the EF path
            if (j < 16) half += d;
            nib[j >> 4] |= (d & 15) << (4 * (j & 15));
            top |= (d >> 4) << (2 * j);
        }
        sec[4 * s0] = base | (half << 48);
        sec[4 * s0 + 1] = nib[0];
        sec[4 * s0 + 2] = nib[1];
        sec[4 * s0 + 3] = top;
    }
    fast->sectors = append(sec.data(), sec.size(), 0);
    fast_built_ = true;
}

void ImageBuilder::read_rank(Cursor& c, DevRank& out) {
    uint64_t nbits = c.pod<uint64_t>();
    uint64_t nw, np, nh;
    const uint8_t* bw = c.vec<uint64_t>(nw);
    c.vec<uint64_t>(np);  // block_rank_pairs: recomputed below in sector form
    c.vec<uint64_t>(nh);  // select hints: never built by the reference (quartet_wtree.cpp:51-53)
    if ((nbits + 63) / 64 > nw) throw FormatError("bit vector shorter than declared");
    auto word = [&](uint64_t i) -> uint64_t {
        if (i >= nw) return 0;
        uint64_t v;
        std::memcpy(&v, bw + 8 * i, 8);
        if ((i + 1) * 64 > nbits) {  // ignore anything past nbits
            uint64_t keep = nbits > i * 64 ? nbits - i * 64 : 0;
            v &= keep >= 64 ? ~uint64_t(0) : ((uint64_t(1) << keep) - 1);
        }
        return v;
    };
    // unit u covers bits [96 u, 96 u + 96); one extra unit so that rank(nbits) needs no special
    // case (rs_bit_vector.hpp:27-29 returns num_ones there)
    if (nbits >> 32) throw FormatError("rank bit vector longer than 2^32 bits");
    auto word32 = [&](uint64_t i) -> uint32_t { return uint32_t(word(i >> 1) >> (32 * (i & 1))); };
    uint64_t nunits = nbits / 96 + 2;
    std::vector<uint32_t> units(nunits * 4, 0);
    uint64_t ones = 0;
    for (uint64_t u = 0; u < nunits; ++u) {
        units[4 * u] = uint32_t(ones);
        for (int j = 0; j < 3; ++j) {
            uint32_t v = word32(3 * u + j);
            units[4 * u + 1 + j] = v;
            ones += uint64_t(__builtin_popcount(v));
        }
    }
    out.units = reinterpret_cast<const uint4*>(append(units.data(), units.size(), 0));
    out.nbits = nbits;
    out.num_ones = ones;
}

void reciprocal96(uint64_t d, uint32_t out[3]) {
    // M = floor((2^96 - 1) / d) + 1 via long division on 32-bit limbs
    unsigned __int128 num = (((unsigned __int128)1) << 96) - 1;
    unsigned __int128 M = num / d + 1;
    out[0] = uint32_t(M);
    out[1] = uint32_t(M >> 32);
    out[2] = uint32_t(M >> 64);
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
    std::vector<uint64_t> hp(ndict);
    for (uint64_t i = 0; i < fd.size; ++i) hp[i] = murmur64_host(fd.get(i), out.seed);
    for (uint64_t i = 0; i < bd.size; ++i) hp[fd.size + i] = murmur64_host(bd.get(i), out.seed);
    out.hashed_pilots = append(hp.data(), ndict, 1);
    out.ranks_are_u16 = ndict <= 65536 ? 1u : 0u;
    auto rank_of_bucket = [&](uint64_t b) -> uint64_t {
        uint64_t r;
        if (b < fr.size) {
            r = fr.get(b);
            if (r >= fd.size) throw FormatError("pilot rank outside dictionary");
        } else {
            r = br.get(b - fr.size);
            if (r >= bd.size) throw FormatError("pilot rank outside dictionary");
            r += fd.size;
        }
        return r;
    };
    if (out.ranks_are_u16) {
        std::vector<uint16_t> rk(nbuckets);
        for (uint64_t b = 0; b < nbuckets; ++b) rk[b] = uint16_t(rank_of_bucket(b));
        out.ranks = append(rk.data(), nbuckets, 8);
    } else {
        if (ndict > 0xFFFFFFFFull) throw FormatError("pilot dictionary too large");
        std::vector<uint32_t> rk(nbuckets);
        for (uint64_t b = 0; b < nbuckets; ++b) rk[b] = uint32_t(rank_of_bucket(b));
        out.ranks = append(rk.data(), nbuckets, 4);
    }
    out.small_divisors =
        (out.table_size < (1ull << 32) && out.dense < (1ull << 32) && out.sparse < (1ull << 32)) ? 1 : 0;
    std::memset(out.m_table, 0, sizeof(out.m_table));
    std::memset(out.m_dense, 0, sizeof(out.m_dense));
    std::memset(out.m_sparse, 0, sizeof(out.m_sparse));
    if (out.small_divisors) {
        reciprocal96(out.table_size, out.m_table);
        if (out.dense) reciprocal96(out.dense, out.m_dense);
        if (out.sparse) reciprocal96(out.sparse, out.m_sparse);
    }
    std::vector<uint64_t> free_vals;
    read_ef(c, out.free_slots, nullptr, &free_vals);
    if (out.free_slots.n != out.table_size - out.num_keys)
        throw FormatError("single_phf: free-slot count mismatch");
    // minimal remap as a plain array: one load instead of an Elias-Fano select
    out.free32 = nullptr;
    bool fits = true;
    for (uint64_t v : free_vals) fits = fits && v < (1ull << 32);
    if (fits) {
        std::vector<uint32_t> f32(free_vals.begin(), free_vals.end());
        out.free32 = append(f32.data(), f32.size(), 4);
        has_free32_.push_back(&out == &img_.minimizer_order ? 0 : 1);
    }
}

void ImageBuilder::parse(const uint8_t* data, uint64_t n, int kmer_bits) {
    if (kmer_bits != 64 && kmer_bits != 128) throw FormatError("kmer_bits must be 64 or 128");
    arena_.clear();
    has_free32_.clear();
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
    img_.maximal_block = uint64_t(img_.w) * img_.n_maximal;
    read_phf(c, img_.minimizer_order);
    read_rank(c, img_.root);
    read_rank(c, img_.left_right);
    read_rank(c, img_.max_none);
    fast_built_ = false;
    read_ef(c, img_.sp, &img_.sp_fast);
    read_phf(c, img_.fallback);
    if (c.p != c.end) throw FormatError("trailing bytes after the .lph image");
    if (img_.minimizer_order.num_keys != img_.distinct_minimizers ||
        img_.root.nbits != img_.distinct_minimizers ||
        img_.left_right.nbits + img_.max_none.nbits != img_.distinct_minimizers)
        throw FormatError("inconsistent minimizer counts");
    if (!(img_.right_start <= img_.none_sizes_start && img_.none_sizes_start <= img_.none_pos_start &&
          img_.none_pos_start < img_.sp.n))
        throw FormatError("inconsistent sizes_and_positions partition");
    fallback_keys_ = img_.fallback.num_keys;
    file_bytes_ = n;
    arena_.resize((arena_.size() + 255) & ~uint64_t(255), 0);
}

void ImageBuilder::rebase_compact(DevCompact& c, const uint8_t* base) {
    c.bits = reinterpret_cast<const uint64_t*>(base + uintptr_t(c.bits));
}
void ImageBuilder::rebase_ef(DevEF& e, const uint8_t* base) {
    e.high = reinterpret_cast<const uint64_t*>(base + uintptr_t(e.high));
    e.block_inv = reinterpret_cast<const int64_t*>(base + uintptr_t(e.block_inv));
    e.sub_inv = reinterpret_cast<const uint16_t*>(base + uintptr_t(e.sub_inv));
    e.overflow = reinterpret_cast<const uint64_t*>(base + uintptr_t(e.overflow));
    rebase_compact(e.low, base);
}
void ImageBuilder::rebase_phf(DevPhf& p, const uint8_t* base, bool has_free32) {
    if (has_free32) p.free32 = reinterpret_cast<const uint32_t*>(base + uintptr_t(p.free32));
    p.ranks = base + uintptr_t(p.ranks);
    p.hashed_pilots = reinterpret_cast<const uint64_t*>(base + uintptr_t(p.hashed_pilots));
    rebase_ef(p.free_slots, base);
}

DevImage ImageBuilder::rebased(const void* device_base) const {
    DevImage d = img_;
    auto* base = static_cast<const uint8_t*>(device_base);
    bool f0 = false, f1 = false;
    for (int w : has_free32_) (w == 0 ? f0 : f1) = true;
    rebase_phf(d.minimizer_order, base, f0);
    rebase_phf(d.fallback, base, f1);
    d.root.units = reinterpret_cast<const uint4*>(base + uintptr_t(d.root.units));
    d.left_right.units = reinterpret_cast<const uint4*>(base + uintptr_t(d.left_right.units));
    d.max_none.units = reinterpret_cast<const uint4*>(base + uintptr_t(d.max_none.units));
    rebase_ef(d.sp, base);
    if (fast_built_) d.sp_fast.sectors = reinterpret_cast<const uint64_t*>(base + uintptr_t(d.sp_fast.sectors));
    return d;
}

}  // namespace lphb
