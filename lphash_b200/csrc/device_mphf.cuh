// Device-side read primitives over the flat image (device_image.h): MurmurHash2-64, exact
// modulo by invariant divisors, PTHash evaluation, rank, Elias-Fano access/pair, and the
// per-super-k-mer probe `mphf::query`.  Every function cites the reference code whose RESULT it
// reproduces bit for bit ("ref" = reference tree, "pthash/" = external/pthash/).
#pragma once
#include <stdint.h>

#include "device_image.h"

namespace lphb {

#define LPHB_DEV __device__ __forceinline__

enum : int { T_LEFT = 0, T_RIGHT = 1, T_MAXIMAL = 2, T_NONE = 3, T_COLLISION = 4 };

// MurmurHash2-64 of one 8-byte word.  ref: pthash/include/utils/hasher.hpp:46-110 (len == 8:
// one block, empty tail); murmurhash2_64::hash(uint64_t) :175-177; default_hash64 :112-114.
LPHB_DEV uint64_t murmur64(uint64_t v, uint64_t seed) {
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

// a % d for 64-bit a and d < 2^32 with M = ceil(2^96 / d) (three 32-bit limbs).
// The reference computes fastmod_u64(a, ceil(2^128/d), d) (pthash/external/fastmod/fastmod.h:
// 56-63, 159-162), which equals a % d exactly; so does this 96-bit variant
// (Lemire-Kaser-Kurz: exact when fraction bits >= 64 + log2 d).
LPHB_DEV uint32_t mod_small(uint64_t a, const uint32_t M[3], uint32_t d) {
    uint32_t a0 = uint32_t(a), a1 = uint32_t(a >> 32);
    uint64_t p00 = uint64_t(M[0]) * a0;
    uint64_t p01 = uint64_t(M[0]) * a1;
    uint64_t p10 = uint64_t(M[1]) * a0;
    uint32_t l0 = uint32_t(p00);
    uint64_t t1 = (p00 >> 32) + uint32_t(p01) + uint32_t(p10);
    uint32_t l1 = uint32_t(t1);
    uint32_t l2 = uint32_t(t1 >> 32) + uint32_t(p01 >> 32) + uint32_t(p10 >> 32) + M[1] * a1 + M[2] * a0;
    uint64_t q = (uint64_t(l0) * d) >> 32;
    q = (uint64_t(l1) * d + q) >> 32;
    q = (uint64_t(l2) * d + q) >> 32;
    return uint32_t(q);
}

// rarely-taken paths are kept out of line so that the hot loops stay small in the instruction cache
#define LPHB_COLD static __device__ __noinline__

LPHB_COLD uint64_t mod_slow(uint64_t a, uint64_t d) { return a % d; }

LPHB_DEV uint64_t mod_any(uint64_t a, const uint32_t M[3], uint64_t d, bool small) {
    return small ? uint64_t(mod_small(a, M, uint32_t(d))) : mod_slow(a, d);
}

// compact_vector::access.  ref: pthash/include/encoders/compact_vector.hpp:229-234 (an unaligned
// 8-byte load there; two aligned words + funnel shift here; identical for width <= 57).
LPHB_DEV uint64_t compact_get(DevCompact const& c, uint64_t i) {
    uint64_t pos = i * c.width;
    const uint64_t* p = c.bits + (pos >> 6);
    uint32_t sh = uint32_t(pos & 63);
    uint64_t lo = __ldg(p), hi = __ldg(p + 1);
    uint64_t v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    return v & c.mask;
}

// position of the r-th (0-based) set bit of x, r < popc(x).  ref: pthash/include/encoders/
// util.hpp:54-97 (select64 via pdep/tzcnt).  Binary search on popcounts.
LPHB_DEV uint32_t select_in_word(uint64_t x, uint32_t r) {
    uint32_t pos = 0;
    uint32_t lo = uint32_t(x), hi = uint32_t(x >> 32);
    uint32_t c = __popc(lo);
    uint32_t v = lo;
    if (r >= c) { r -= c; v = hi; pos = 32; }
    c = __popc(v & 0xFFFFu);
    if (r >= c) { r -= c; v >>= 16; pos += 16; }
    c = __popc(v & 0xFFu);
    if (r >= c) { r -= c; v >>= 8; pos += 8; }
    c = __popc(v & 0xFu);
    if (r >= c) { r -= c; v >>= 4; pos += 4; }
    c = __popc(v & 0x3u);
    if (r >= c) { r -= c; v >>= 2; pos += 2; }
    if (r >= (v & 1u)) pos += 1;
    return pos;
}

// darray1::select.  ref: pthash/include/encoders/darray.hpp:51-76 (block 1024, subblock 32).
LPHB_COLD uint64_t darray_select(DevEF const& e, uint64_t idx) {
    int64_t bp = __ldg(e.block_inv + (idx >> 10));
    if (bp < 0) return __ldg(e.overflow + uint64_t(-bp - 1) + (idx & 1023));
    uint64_t start = uint64_t(bp) + __ldg(e.sub_inv + (idx >> 5));
    uint32_t rem = uint32_t(idx & 31);
    if (rem == 0) return start;
    uint64_t wd = start >> 6;
    uint64_t cur = __ldg(e.high + wd) & (~uint64_t(0) << (start & 63));
    for (;;) {
        uint32_t pc = uint32_t(__popcll(cur));
        if (rem < pc) break;
        rem -= pc;
        cur = __ldg(e.high + (++wd));
    }
    return (wd << 6) + select_in_word(cur, rem);
}

// ef_sequence::access.  ref: include/ef_sequence.hpp:77-81; pthash ef_sequence.hpp:55-59.
LPHB_DEV uint64_t ef_access(DevEF const& e, uint64_t i) {
    uint64_t hi = darray_select(e, i) - i;
    return e.low.width ? ((hi << e.low.width) | compact_get(e.low, i)) : hi;
}

// ef_sequence::pair (values i and i+1).  ref: include/ef_sequence.hpp:83-94; the successor bit
// is found as pthash bit_vector::unary_iterator(pos+1).next() does (bit_vector.hpp:235-258).
LPHB_DEV void ef_pair(DevEF const& e, uint64_t i, uint64_t& v1, uint64_t& v2) {
    uint64_t pos = darray_select(e, i);
    uint64_t q = pos + 1;
    uint64_t wd = q >> 6;
    uint64_t cur = __ldg(e.high + wd) & (~uint64_t(0) << (q & 63));
    while (cur == 0) cur = __ldg(e.high + (++wd));
    uint64_t nxt = (wd << 6) + uint64_t(__ffsll((long long)cur) - 1);
    uint32_t l = e.low.width;
    uint64_t h1 = pos - i, h2 = nxt - i - 1;
    if (l) {
        v1 = (h1 << l) | compact_get(e.low, i);
        v2 = (h2 << l) | compact_get(e.low, i + 1);
    } else {
        v1 = h1;
        v2 = h2;
    }
}

// S[i] and d = S[i+1] - S[i] from the prefix-sector layout (device_image.h: DevPrefix): one
// 32-byte sector, SWAR sums of 4-bit / 2-bit fields.  Same values as ef_sequence::pair(i)
// (ref: include/ef_sequence.hpp:83-94).
LPHB_DEV void prefix_pair(DevPrefix const& t, uint64_t i, uint64_t& v1, uint32_t& d) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(t.sectors + 4 * (i >> 5));
    const ulonglong2 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t j = uint32_t(i) & 31u, t16 = j & 15u;
    const bool upper = j >= 16u;
    const uint64_t nib = upper ? b.x : a.y;
    const uint32_t top = upper ? uint32_t(b.y >> 32) : uint32_t(b.y);
    d = (uint32_t(nib >> (4 * t16)) & 15u) | (((top >> (2 * t16)) & 3u) << 4);
    // sum of the first t16 deltas of this half
    const uint64_t nm = nib & ~(~uint64_t(0) << (4 * t16));
    const uint32_t tm = top & ~(~0u << (2 * t16));
    const uint32_t lo = uint32_t(nm), hi = uint32_t(nm >> 32);
    uint32_t sn = (lo & 0x0F0F0F0Fu) + ((lo >> 4) & 0x0F0F0F0Fu) + (hi & 0x0F0F0F0Fu) + ((hi >> 4) & 0x0F0F0F0Fu);
    sn = (sn * 0x01010101u) >> 24;  // <= 15 * 15
    uint32_t sp = (tm & 0x33333333u) + ((tm >> 2) & 0x33333333u);
    sp = (sp + (sp >> 4)) & 0x0F0F0F0Fu;
    sp = (sp * 0x01010101u) >> 24;  // <= 15 * 3
    const uint32_t half = upper ? uint32_t(a.x >> 48) : 0u;
    v1 = (a.x & 0x0000FFFFFFFFFFFFull) + half + sn + 16u * sp;
}

// bit `pos` and the number of ones before it: one 16-byte unit {ones before, 96 bits}.  Equals
// rs_bit_vector::operator[] and rank (ref: include/rs_bit_vector.hpp:27-38, 101-114; pos == nbits
// is served by the terminal unit).
LPHB_DEV void rank_unit(DevRank const& r, uint32_t pos, uint32_t& bit, uint32_t& ones_before) {
    const uint32_t u = __umulhi(pos, 0xAAAAAAABu) >> 6;  // pos / 96
    const uint32_t off = pos - u * 96u;
    const uint4 v = __ldg(r.units + u);
    const uint32_t wi = off >> 5, sh = off & 31u;
    const uint32_t cur = wi == 0 ? v.y : (wi == 1 ? v.z : v.w);
    uint32_t ones = v.x + __popc(cur & ((1u << sh) - 1u));
    if (wi > 0) ones += __popc(v.y);
    if (wi > 1) ones += __popc(v.z);
    ones_before = ones;
    bit = (cur >> sh) & 1u;
}

// quartet_wtree::rank_of.  ref: src/quartet_wtree.cpp:84-99.
LPHB_DEV void wtree_rank_of(DevImage const& f, uint64_t idx, uint32_t& type, uint64_t& rank) {
    uint32_t msb, lsb, ones;
    rank_unit(f.root, uint32_t(idx), msb, ones);
    const uint32_t r = msb ? ones : uint32_t(idx) - ones;
    rank_unit(msb ? f.max_none : f.left_right, r, lsb, ones);
    rank = lsb ? ones : r - ones;
    type = (msb << 1) | lsb;
}

// pthash::single_phf::position.  ref: pthash/include/single_phf.hpp:55-65;
// skew_bucketer::bucket pthash/include/utils/bucketers.hpp:17-22 (T = (uint64_t)(0.6*UINT64_MAX)
// evaluated in double = 0x9999999999999800); dual/dictionary access encoders.hpp:167-170,268-271.
// position before the minimal remap (may be >= num_keys: then free_slots resolves it)
LPHB_DEV uint64_t phf_raw_position(DevPhf const& p, uint64_t h) {
    const uint64_t T = 0x9999999999999800ULL;
    bool sm = p.small_divisors != 0;
    uint64_t b = h < T ? mod_any(h, p.m_dense, p.dense, sm)
                       : p.dense + mod_any(h, p.m_sparse, p.sparse, sm);
    uint32_t rk = p.ranks_are_u16 ? uint32_t(__ldg(reinterpret_cast<const uint16_t*>(p.ranks) + b))
                                  : __ldg(reinterpret_cast<const uint32_t*>(p.ranks) + b);
    uint64_t hp = __ldg(p.hashed_pilots + rk);
    return mod_any(h ^ hp, p.m_table, p.table_size, sm);
}

LPHB_DEV uint64_t phf_position(DevPhf const& p, uint64_t h) {
    uint64_t pos = phf_raw_position(p, h);
    if (pos < p.num_keys) return pos;
    if (p.free32) return __ldg(p.free32 + (pos - p.num_keys));
    return ef_access(p.free_slots, pos - p.num_keys);  // cold: darray_select is out of line
}

// fallback_kmer_order(kmer).  ref: include/constants.hpp:56-70 (fallback_hasher: 64-bit kmer_t
// hashes the single word; 128-bit kmer_t xors murmur(lo, seed) with murmur(hi, ~seed)).
LPHB_DEV uint64_t fallback_order(DevImage const& f, uint64_t lo, uint64_t hi) {
    uint64_t h = murmur64(lo, f.fallback.seed);
    if (f.kmer_bits == 128) h ^= murmur64(hi, ~f.fallback.seed);
    return phf_position(f.fallback, h);
}

// What one super-k-mer probe yields: hval of a k-mer whose minimizer sits at offset p is
// base + slope * p (mod 2^64) for the four regular types; colliding minimizers send every k-mer
// through fallback_order instead (hval = base + fallback).
struct Probe {
    uint64_t base;
    int32_t slope;  // +1: LEFT, MAXIMAL;  -1: RIGHT, NONE;  0: COLLISION
    uint32_t type;
};

// mphf::query without the k-mer-dependent part.  ref: src/partitioned_mphf.cpp:292-339.
//   LEFT     g = EF[rank] + w*n_max,                  l = p
//   RIGHT    (v1,v2) = EF.pair(right_start+rank), v2>v1: g = v1 + w*n_max, l = (k-m) - p
//   COLL     v2 == v1: g = EF[none_pos_start] + w*n_max, l = fallback(kmer)
//   MAXIMAL  g = w*rank,                              l = p
//   NONE     g = EF[none_sizes_start+rank] + w*n_max, l = EF.diff(none_pos_start+rank) - p
// file-layout Elias-Fano path (only when the prefix sectors could not be built)
LPHB_COLD Probe probe_bucket_ef(DevImage const& f, uint32_t type, uint64_t rk) {
    Probe out;
    if (type == T_MAXIMAL) {
        out.base = uint64_t(f.w) * rk;
        out.slope = 1;
        out.type = T_MAXIMAL;
    } else if (type == T_LEFT) {
        out.base = ef_access(f.sp, rk) + f.maximal_block;
        out.slope = 1;
        out.type = T_LEFT;
    } else if (type == T_RIGHT) {
        uint64_t v1, v2;
        ef_pair(f.sp, f.right_start + rk, v1, v2);
        if (v2 == v1) {
            out.base = f.collision_base;
            out.slope = 0;
            out.type = T_COLLISION;
        } else {
            out.base = v1 + f.maximal_block + uint64_t(f.k - f.m);
            out.slope = -1;
            out.type = T_RIGHT;
        }
    } else {
        uint64_t v1, v2;
        ef_pair(f.sp, f.none_pos_start + rk, v1, v2);
        out.base = ef_access(f.sp, f.none_sizes_start + rk) + f.maximal_block + (v2 - v1);
        out.slope = -1;
        out.type = T_NONE;
    }
    return out;
}

LPHB_DEV Probe probe_bucket(DevImage const& f, uint64_t bucket) {
    Probe out;
    uint32_t type;
    uint64_t rk;
    wtree_rank_of(f, bucket, type, rk);
    if (f.sp_fast.sectors) {
        // one prefix sector for every non-MAXIMAL type, a second one (offset only) for NONE
        const uint64_t ia = type == T_LEFT ? rk : (type == T_RIGHT ? f.right_start + rk : f.none_sizes_start + rk);
        uint64_t v1 = 0;
        uint32_t d = 0, d2 = 0;
        if (type != T_MAXIMAL) prefix_pair(f.sp_fast, ia, v1, d);
        if (type == T_NONE) {
            uint64_t unused;
            prefix_pair(f.sp_fast, f.none_pos_start + rk, unused, d2);
        }
        if (type == T_MAXIMAL) {
            out.base = uint64_t(f.w) * rk;
            out.slope = 1;
            out.type = T_MAXIMAL;
        } else if (type == T_LEFT) {
            out.base = v1 + f.maximal_block;
            out.slope = 1;
            out.type = T_LEFT;
        } else if (type == T_RIGHT) {
            const bool coll = d == 0;
            out.base = coll ? f.collision_base : v1 + f.maximal_block + uint64_t(f.k - f.m);
            out.slope = coll ? 0 : -1;
            out.type = coll ? T_COLLISION : T_RIGHT;
        } else {
            out.base = v1 + f.maximal_block + d2;
            out.slope = -1;
            out.type = T_NONE;
        }
        return out;
    }
    return probe_bucket_ef(f, type, rk);
}

LPHB_DEV Probe probe_minimizer(DevImage const& f, uint64_t minimizer) {
    return probe_bucket(f, phf_position(f.minimizer_order, murmur64(minimizer, f.minimizer_order.seed)));
}

LPHB_DEV uint64_t probe_hval(Probe const& pr, uint32_t p) {
    return pr.slope > 0 ? pr.base + p : pr.base - p;
}

// ASCII -> 2-bit code, 4 = invalid.  ref: src/constants.cpp:5-13.
LPHB_DEV uint32_t nt4(uint32_t ch) {
    uint32_t u = ch & 0xDFu;  // fold case
    uint32_t code = ((ch >> 1) ^ (ch >> 2)) & 3u;  // A0 C1 G2 T/U3
    bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T') | (u == 'U');
    return ok ? code : 4u;
}

}  // namespace lphb
