// Tiled streaming-query kernel (the hot path) for window widths W = k-m+1 <= 17.
//
// The concatenated base stream is cut into tiles of TILE = 8 warps x 31 lanes x 16 k-mer starts.
// One CTA (256 threads) per tile, five phases separated by CTA barriers:
//
//  A  load+pack   one coalesced 16-byte load per thread (16 ASCII bases) -> one 32-bit word of
//                 2-bit codes (first base in the most significant bits, the reference's m-mer /
//                 k-mer orientation, partitioned_mphf.hpp:106-108) -> shared memory.  Non-ACGT
//                 bytes flag their contig dirty (it is then recomputed by the exact sequential
//                 kernel).  Warp 0 meanwhile rasterises contig seams into a
//                 bitmask of invalid k-mer starts.
//  B  scan        each thread owns 16 consecutive k-mer starts.  16 m-mer hashes from registers
//                 (MurmurHash2-64, seeded); each is reduced to a 32-bit key = top 27 bits of the
//                 hash | 5-bit thread-local position, so that ONE unsigned min picks the smaller
//                 hash and, between equal keys, the leftmost.  The W-1 keys a thread lacks come
//                 from lane+1 by warp shuffle (lane 31 only feeds lane 30: warps overlap by one
//                 lane).  Sliding minimum = sparse table of 3-input minima (VIMNMX3): spans of 3,
//                 9, then two spans cover the window.  A second pass with the position bits
//                 complemented finds the RIGHTMOST minimum; if the two differ anywhere the 27-bit
//                 keys tied (true repeat or truncation tie) and that thread recomputes its 16
//                 windows from the full 64-bit hashes (out of line, rare).  Result: minimizer
//                 offset of every k-mer + mask of positions that are some k-mer's minimizer.
//  C  compact     minimizer positions of the warp -> dense list in shared memory (warp scan).
//  D  probe       one lane per distinct minimizer position: m-mer -> PTHash -> bucket table
//                 (device_mphf.cuh) -> {B, ns} with  code(k-mer at q) = B + ns * q.
//  E  emit        position-parallel: lane l handles k-mers l, l+32, ...: two byte loads find the
//                 entry, one IMAD.WIDE makes the code, 8-byte stores are fully coalesced.
//                 K-mers of colliding minimizers are queued and resolved through
//                 fallback_kmer_order (partitioned_mphf.cpp:308-313).
//
// Every k-mer's code is a pure function of its own k bases, so tiles only share
// k-1 bases of read overlap and nothing else.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_mphf.cuh"
#include "query_kernels.cuh"

namespace lphb {

namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kS = 16;                         // k-mer starts per thread = bases per packed word
constexpr int kLanes = 31;                     // producing lanes per warp (lane 31 only feeds lane 30)
constexpr int kStrip = kLanes * kS;            // 496 k-mer starts per warp pass
constexpr int kStrips = 2;                     // passes per warp
constexpr int kWarpKmers = kStrips * kStrip;   // 992 k-mer starts per warp
constexpr int kTile = kWarps * kWarpKmers;     // 3968 k-mer starts per tile (CTA)
constexpr int kMaskWords = kTile / 32;         // 124
constexpr int kMaskSlots = kMaskWords + 1;
constexpr int kWarpSlots = kWarpKmers + 32;    // positions a minimizer of the warp's k-mers can sit at
constexpr int kWarpMaskWords = kWarpSlots / 32;  // 32: one word per lane
constexpr int kProbes = 6;                     // probes a lane keeps in flight
constexpr int kCap = 32 * kProbes;             // probe results held at once per warp (index fits a byte)
struct Entry {                                 // code of the k-mer at warp-local q = B + ns * q (mod 2^64)
    uint32_t lo;                               // low word of B (the high word lives in s_hi)
    int32_t ns;                                // -1: LEFT/MAXIMAL, +1: RIGHT/NONE, 0: colliding minimizer
};
constexpr int kWarpBytes = kCap * (8 + 4) + kWarpSlots * (1 + 1 + 2) + (kWarpMaskWords + 4 + 16) * 4;
constexpr int kPackedSlots = 256;
constexpr int kSmemBytes = kWarps * kWarpBytes + kPackedSlots * 4 + kMaskSlots * (4 + 2) + 16;
static_assert(kWarpMaskWords == 32 && kWarpBytes % 16 == 0, "per-warp layout");

struct TileArgs {
    const char* abase;          // 16-byte aligned; stream position pos0 lives here
    int64_t pos0;               // stream position of abase (may be < first_base by < 16)
    const uint32_t* tile_c0;    // per tile: contig containing the tile's first in-range position
    const uint64_t* tile_out;   // per tile: number of valid k-mer starts before it
    uint32_t n_tiles;
};

// 4 ASCII bytes -> 8 bits of 2-bit codes, byte 0 in the top 2 bits.  (x>>1 ^ x>>2) & 3 maps
// A,a->0 C,c->1 G,g->2 T,t,U,u->3 (src/constants.cpp:5-13 for the valid bytes).
__device__ __forceinline__ uint32_t codes4(uint32_t x) { return ((x >> 1) ^ (x >> 2)) & 0x03030303u; }
__device__ __forceinline__ uint32_t pack4(uint32_t y) { return (y * 0x40100401u) >> 24; }
// nonzero iff one of the 4 bytes is not in {A,C,G,T,U,a,c,g,t,u}: rebuild the canonical upper-case
// letter of each code with a byte permute and compare (T and U both map to 3: two tables).
__device__ __forceinline__ uint32_t bad4(uint32_t x, uint32_t y) {
    uint32_t z = y | (y >> 4);
    uint32_t sel = __byte_perm(z, 0u, 0x4420u);  // nibble i = code of byte i
    uint32_t c1 = __byte_perm(0x54474341u, 0u, sel);  // "ACGT"[code]
    uint32_t c2 = __byte_perm(0x55474341u, 0u, sel);  // "ACGU"[code]
    uint32_t u = x & 0xDFDFDFDFu;
    return (u ^ c1) & (u ^ c2);
}

// flag the contig of every in-range non-ACGT byte of a 16-byte word (rare; out of line)
static __device__ __noinline__ void mark_dirty(DevBatch const& b, uint4 v, int64_t wpos) {
    const int64_t first = int64_t(b.first_base), end = int64_t(b.end_base);
    uint32_t xs[4] = {v.x, v.y, v.z, v.w};
    for (int j = 0; j < 16; ++j) {
        int64_t p = wpos + j;
        uint32_t ch = (xs[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (p >= first && p < end && nt4(ch) > 3) {
            uint64_t lo = 0, hi = b.n_contigs;
            while (hi - lo > 1) {
                uint64_t mid = (lo + hi) >> 1;
                if (int64_t(__ldg(b.offsets + mid)) <= p) lo = mid; else hi = mid;
            }
            b.dirty[lo] = 1;
        }
    }
}

// code of a k-mer whose minimizer collides: fallback_kmer_order (rare; out of line)
static __device__ __noinline__ uint64_t fallback_code(DevImage const& f, uint64_t klo, uint64_t khi) {
    return f.collision_base + fallback_order(f, klo, khi);
}

template <int K, int M>
struct Cfg {
    static constexpr int W = K - M + 1;
    static constexpr int NW = (kS + K - 1 + 15) / 16;        // packed words a thread reads
    static constexpr int NH = kS + W - 1;                    // m-mers under a thread's 16 windows
    static constexpr int TileWords = kTile / 16 + NW;        // words staged per tile
    static_assert(W >= 1 && W <= 17, "tiled kernel: window must fit one shuffle hop");
    static_assert(NH <= 32, "thread-local minimizer positions must fit the 5-bit key field");
    static_assert(M <= 31 && K <= 63, "k, m out of range");
    static_assert(TileWords <= kPackedSlots, "packed tile must fit its shared-memory array");
};

// a * 0xc6a4a7935bd1e995 mod 2^64 in three multiply-adds (IMAD.WIDE + 2 IMAD, all on the FMA pipe)
__device__ __forceinline__ uint64_t mul_murmur(uint32_t lo, uint32_t hi) {
    uint64_t w = uint64_t(lo) * 0x5bd1e995u;
    uint32_t h = uint32_t(w >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(lo), "r"(0xc6a4a793u));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(hi), "r"(0x5bd1e995u));
    return (uint64_t(h) << 32) | uint32_t(w);
}
__device__ __forceinline__ uint64_t mul_murmur(uint64_t a) { return mul_murmur(uint32_t(a), uint32_t(a >> 32)); }

// 16 bases starting at base t (compile-time after unrolling) of a thread's packed words
template <int NW>
__device__ __forceinline__ uint32_t win16(const uint32_t (&wds)[NW], int t) {
    const int q = t >> 4, r = (t & 15) * 2;
    if (r == 0) return wds[q];
    const uint32_t nxt = q + 1 < NW ? wds[q + 1] : 0u;
    return __funnelshift_l(nxt, wds[q], r);
}

// out[i] = min(k[i .. i+W-1]) for i < 16: sparse table of 3-input minima (spans 3, 9), the window
// is then covered by two (overlapping) spans.  Unused table entries vanish at compile time.
template <int W, int NH>
__device__ __forceinline__ void window_min(const uint32_t (&k)[NH], uint32_t (&out)[kS]) {
    if constexpr (W == 1) {
#pragma unroll
        for (int i = 0; i < kS; ++i) out[i] = k[i];
    } else if constexpr (W == 2) {
#pragma unroll
        for (int i = 0; i < kS; ++i) out[i] = min(k[i], k[i + 1]);
    } else {
        uint32_t m3[NH];
#pragma unroll
        for (int i = 0; i + 2 < NH; ++i) m3[i] = __vimin3_u32(k[i], k[i + 1], k[i + 2]);
        if constexpr (W == 3) {
#pragma unroll
            for (int i = 0; i < kS; ++i) out[i] = m3[i];
        } else if constexpr (W <= 6) {
#pragma unroll
            for (int i = 0; i < kS; ++i) out[i] = min(m3[i], m3[i + W - 3]);
        } else if constexpr (W < 9) {
#pragma unroll
            for (int i = 0; i < kS; ++i) out[i] = __vimin3_u32(m3[i], m3[i + 3], m3[i + W - 3]);
        } else {
            uint32_t m9[NH];
#pragma unroll
            for (int i = 0; i + 8 < NH; ++i) m9[i] = __vimin3_u32(m3[i], m3[i + 3], m3[i + 6]);
            if constexpr (W == 9) {
#pragma unroll
                for (int i = 0; i < kS; ++i) out[i] = m9[i];
            } else if constexpr (W <= 12) {
#pragma unroll
                for (int i = 0; i < kS; ++i) out[i] = min(m9[i], m3[i + W - 3]);
            } else {
#pragma unroll
                for (int i = 0; i < kS; ++i) out[i] = min(m9[i], m9[i + W - 9]);
            }
        }
    }
}

// m-mer starting at tile-local base g, from the packed tile in shared memory
template <int M>
__device__ __forceinline__ uint64_t mmer_at(const uint32_t* s_packed, int g) {
    const int wi = g >> 4, r = (g & 15) * 2;
    const uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1], w2 = s_packed[wi + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, r), lo = __funnelshift_l(w2, w1, r);
    return ((uint64_t(hi) << 32) | lo) >> (64 - 2 * M);
}

// Exact minimizer offsets of the 16 k-mers starting at tile-local base g0, from the full 64-bit
// hashes (strict '<' keeps the leftmost on ties: partitioned_mphf.hpp:124,152,159).  Taken only by
// threads whose 27-bit keys tied; out of line.  Returns the mask of minimizer positions.
template <int K, int M>
static __device__ __noinline__ uint32_t exact_strip(const uint32_t* s_packed, int g0, uint64_t seed,
                                                    uint8_t* pos_out) {
    constexpr int W = K - M + 1, NH = kS + W - 1;
    uint64_t h[NH];
#pragma unroll 1
    for (int j = 0; j < NH; ++j) h[j] = murmur64(mmer_at<M>(s_packed, g0 + j), seed);
    uint32_t marks = 0;
#pragma unroll 1
    for (int i = 0; i < kS; ++i) {
        int best = i;
        for (int j = i + 1; j < i + W; ++j)
            if (h[j] < h[best]) best = j;
        pos_out[i] = uint8_t(best);
        marks |= 1u << best;
    }
    return marks;
}

// Codes of the k-mers of a warp, position-parallel: lane l handles k-mers l, l+32, ...; two byte
// loads find the entry of the k-mer's minimizer, code = B + ns * q, coalesced 8-byte stores.
//
// Plain form, for a warp whose 992 starts all yield a code, whose entries share the high word
// `hi` of B and cannot carry out of the low word (1024 <= lo < 2^32 - 1024), with no colliding
// minimizer and a single chunk: one 32-bit multiply-add per code.
__device__ __forceinline__ void emit_plain(int lane, const uint8_t* s_pos, const uint8_t* s_ref,
                                           const Entry* s_ent, uint32_t hi, uint64_t* out_warp) {
    const uint8_t* pos_l = s_pos + lane;
    const uint8_t* ref_l = s_ref + (lane & 16);
    uint2* o = reinterpret_cast<uint2*>(out_warp + lane);
#pragma unroll 8
    for (int r = 0; r < kWarpKmers / 32; ++r) {
        const int mp = int(pos_l[r * 32]) + r * 32;  // (+ lane & 16) warp-local position of the minimizer
        const int2 e = *reinterpret_cast<const int2*>(s_ent + ref_l[mp]);
        uint32_t lo;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(lo) : "r"(e.y), "r"(lane + r * 32), "r"(e.x));
        __stcs(o + r * 32, make_uint2(lo, hi));
    }
}

// General form.  Per group of 32 starts the (warp-uniform) word of the invalid-start mask tells
// whether starts without a code (contig seams) must be skipped and the output index compacted;
// k-mers of colliding minimizers are queued in s_list; kChunked (more than kCap minimizers):
// entries outside [i0, i1) are left to their own chunk.
template <bool kChunked>
__device__ __forceinline__ void emit_general(int lane, int wbase, uint32_t i0, uint32_t i1,
                                             const uint8_t* s_pos, const uint8_t* s_ref,
                                             const uint32_t* s_minmask, const uint16_t* s_wpre,
                                             const Entry* s_ent, const uint32_t* s_hi,
                                             uint16_t* s_list, uint32_t* s_n_fb,
                                             const uint32_t* s_invalid, const uint16_t* s_invpre,
                                             uint64_t* out) {
    const int lane16 = lane & 16;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t* inv = s_invalid + (wbase >> 5);  // wbase is a multiple of 32
    const uint16_t* pre = s_invpre + (wbase >> 5);
    uint64_t* out_l = out + wbase + lane;
#pragma unroll 2
    for (int r = 0; r < kWarpKmers / 32; ++r) {
        const int q = lane + r * 32;
        const uint32_t mw = inv[r];  // uniform in the warp
        if ((mw >> lane) & 1u) continue;
        const int mp = int(s_pos[q]) + lane16 + r * 32;  // warp-local position of q's minimizer
        uint32_t idx;
        if (kChunked) {  // global list index of position mp: rank among the marked positions
            idx = s_wpre[mp >> 5] + __popc(s_minmask[mp >> 5] & ((1u << (mp & 31)) - 1u));
            if (idx < i0 || idx >= i1) continue;
            idx -= i0;
        } else {
            idx = s_ref[mp];
        }
        const Entry e = s_ent[idx];
        if (e.ns == 0) {  // colliding minimizer: needs the k-mer itself
            s_list[atomicAdd(s_n_fb, 1u)] = uint16_t(q);
            continue;
        }
        const uint64_t B = (uint64_t(s_hi[idx]) << 32) | e.lo;
        const uint64_t code = B + uint64_t(int64_t(e.ns) * int64_t(q));
        __stcs(out_l + (r * 32 - int(pre[r]) - __popc(mw & lt)), code);
    }
}

template <int K, int M>
__global__ void __launch_bounds__(kThreads, 6)
k_query_tiled(const __grid_constant__ DevImage f, const __grid_constant__ DevBatch b,
              const __grid_constant__ TileArgs a) {
    using C = Cfg<K, M>;
    constexpr int W = C::W, NW = C::NW, NH = C::NH;

    // dynamic shared memory: one private region per warp (warp-local coordinates) + tile-wide
    // packed bases and invalid-start bitmask
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* mine = smem_raw + warp * kWarpBytes;
    Entry* s_ent = reinterpret_cast<Entry*>(mine);                          // per probe (list index)
    uint32_t* s_hi = reinterpret_cast<uint32_t*>(s_ent + kCap);            // per probe: high word of B
    uint8_t* s_pos = reinterpret_cast<uint8_t*>(s_hi + kCap);              // per k-mer: thread-local minimizer position
    uint8_t* s_ref = s_pos + kWarpSlots;                                   // per position: chunk-local list index
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_ref + kWarpSlots);    // minimizer positions (chunk); later colliding k-mers
    uint32_t* s_minmask = reinterpret_cast<uint32_t*>(s_list + kWarpSlots);  // bit b: position b is some k-mer's minimizer
    uint32_t* s_n_fb = s_minmask + kWarpMaskWords;                         // colliding k-mers queued
    uint16_t* s_wpre = reinterpret_cast<uint16_t*>(s_n_fb + 4);            // per mask word: marked positions before it
    uint32_t* s_packed = reinterpret_cast<uint32_t*>(smem_raw + kWarps * kWarpBytes);  // 2-bit bases (tile)
    uint32_t* s_invalid = s_packed + kPackedSlots;     // bit q: k-mer start q produces no code (tile)
    uint16_t* s_invpre = reinterpret_cast<uint16_t*>(s_invalid + kMaskSlots);  // invalid starts before word

    const uint32_t tile = blockIdx.x;
    const int64_t T0 = a.pos0 + int64_t(tile) * kTile;  // stream position of tile-local 0
    const int64_t first = int64_t(b.first_base), end = int64_t(b.end_base);

    // ---------------------------------------------------------------- A: load + pack ------------
    if (tid < kMaskSlots) s_invalid[tid] = 0;
    s_minmask[lane] = 0;
    if (lane == 0) *s_n_fb = 0;
    for (int t = tid; t < kPackedSlots; t += kThreads) {
        int64_t wpos = T0 + int64_t(t) * 16;
        uint32_t word = 0;
        if (t < C::TileWords && wpos + 16 > first && wpos < end) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(a.abase + (wpos - a.pos0)));
            uint32_t y0 = codes4(v.x), y1 = codes4(v.y), y2 = codes4(v.z), y3 = codes4(v.w);
            word = (pack4(y0) << 24) | (pack4(y1) << 16) | (pack4(y2) << 8) | pack4(y3);
            uint32_t bad = bad4(v.x, y0) | bad4(v.y, y1) | bad4(v.z, y2) | bad4(v.w, y3);
            if (bad) mark_dirty(b, v, wpos);  // rare
        }
        s_packed[t] = word;
    }
    __syncthreads();  // zeroed masks + packed words visible

    // warp 0: rasterise the k-mer starts that produce no code (contig seams, short contigs,
    // positions outside [first, end)) into s_invalid
    if (warp == 0) {
        const int64_t tile_end = T0 + kTile;
        if (T0 < first) {  // head padding of the first tile (< 16 positions)
            int n = int(first - T0);
            if (lane == 0) atomicOr(&s_invalid[0], (1u << n) - 1u);
        }
        uint64_t c = a.tile_c0[tile];
        for (;; c += 32) {
            uint64_t cc = c + lane;
            bool live = cc < b.n_contigs;
            int64_t s = live ? int64_t(__ldg(b.offsets + cc)) : end;
            int64_t e = live ? int64_t(__ldg(b.offsets + cc + 1)) : end;
            if (live && s < tile_end) {
                // starts in [max(e-K+1, s), e) have fewer than K bases left in their contig
                int64_t lo = e - (K - 1) > s ? e - (K - 1) : s;
                int64_t hi = e;
                if (lo < T0) lo = T0;
                if (hi > tile_end) hi = tile_end;
                for (int64_t q = lo; q < hi;) {
                    int ql = int(q - T0);
                    int wbit = ql & 31;
                    int n = int(hi - q) < 32 - wbit ? int(hi - q) : 32 - wbit;
                    uint32_t bits = (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)) << wbit;
                    atomicOr(&s_invalid[ql >> 5], bits);
                    q += n;
                }
            }
            // go on while the contig after lane 31's also starts inside the tile
            if (!__any_sync(0xFFFFFFFFu, lane == 31 && live && e < tile_end)) break;
        }
        if (end < tile_end) {  // past the last base of the batch
            int lo = end > T0 ? int(end - T0) : 0;
            for (int wd = (lo >> 5) + lane; wd < kMaskWords; wd += 32) {
                uint32_t bits = 0xFFFFFFFFu;
                if (wd == (lo >> 5)) bits <<= (lo & 31);
                atomicOr(&s_invalid[wd], bits);
            }
        }
        __syncwarp();
        // exclusive prefix of invalid counts per mask word (124 words: 4 per lane)
        uint32_t cnt[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int wd = lane * 4 + j;
            cnt[j] = wd < kMaskWords ? __popc(s_invalid[wd]) : 0;
            sum += cnt[j];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += v;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int wd = lane * 4 + j;
            if (wd < kMaskSlots) s_invpre[wd] = uint16_t(run);
            run += cnt[j];
        }
    }
    __syncthreads();  // invalid-start mask ready; from here on every warp runs on its own

    // ---------------------------------------------------------------- B: per-thread scan --------
    const int wbase = warp * kWarpKmers;  // tile-local position of this warp's first k-mer
    const uint64_t h0 = f.mm_seed ^ (8 * kMurmurM);
    uint32_t keymask;  // ~31 held in a register so that (hash & ~31) | position is one LOP3
    asm volatile("mov.u32 %0, 0xFFFFFFE0;" : "=r"(keymask));
#pragma unroll 1
    for (int strip = 0; strip < kStrips; ++strip) {
        const int lseg = strip * kStrip + lane * kS;  // warp-local position of this thread's first k-mer
        uint32_t wds[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) wds[j] = s_packed[((wbase + lseg) >> 4) + j];

        // keys: top 27 bits of the m-mer's hash | thread-local position (own: 0..15, the W-1
        // received from lane+1: 16..)
        uint32_t key[NH];
#pragma unroll
        for (int j = 0; j < kS; ++j) {
            uint32_t v_lo, v_hi;
            if constexpr (M <= 16) {
                v_lo = M == 16 ? win16<NW>(wds, j) : win16<NW>(wds, j) >> (32 - 2 * M);
                v_hi = 0;
            } else {
                v_lo = win16<NW>(wds, j + M - 16);
                v_hi = win16<NW>(wds, j) >> (64 - 2 * M);
            }
            // MurmurHash2-64 (device_mphf.cuh: murmur64) up to its last multiply: the final
            // h ^= h >> 47 cannot change the top 32 bits
            uint64_t x = M <= 16 ? uint64_t(v_lo) * kMurmurM : mul_murmur(v_lo, v_hi);
            x ^= x >> 47;
            x = mul_murmur(x);
            uint64_t h = mul_murmur(h0 ^ x);
            h ^= h >> 47;
            h = mul_murmur(h);
            asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(key[j]) : "r"(uint32_t(h >> 32)), "r"(keymask), "r"(uint32_t(j)));  // (h & mask) | j
        }
#pragma unroll
        for (int j = 0; j < W - 1; ++j) key[kS + j] = __shfl_down_sync(0xFFFFFFFFu, key[j] + 16u, 1);

        uint32_t mn[kS];
        window_min<W, NH>(key, mn);
        // same windows with the position bits complemented: the minimum is now the RIGHTMOST one
        // among equal 27-bit keys; both agree on every window <=> no two candidates tied
        uint32_t agree = 31u;
        {
            uint32_t rkey[NH], rmn[kS];
#pragma unroll
            for (int j = 0; j < NH; ++j) rkey[j] = key[j] ^ 31u;
            window_min<W, NH>(rkey, rmn);
#pragma unroll
            for (int i = 0; i < kS; ++i) agree &= mn[i] ^ rmn[i];
        }
        if (lane < kLanes) {  // lane 31 only feeds keys to lane 30
            uint32_t marks = 0;  // bit j: thread-local position j is the minimizer of one of my k-mers
            if (agree == 31u) {
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < kS; ++i) marks |= __funnelshift_l(0u, 1u, mn[i]);  // 1 << (mn & 31)
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    uint32_t t0 = __byte_perm(mn[4 * g4], mn[4 * g4 + 1], 0x0040);
                    uint32_t t1 = __byte_perm(mn[4 * g4 + 2], mn[4 * g4 + 3], 0x0040);
                    pk[g4] = __byte_perm(t0, t1, 0x5410) & 0x1F1F1F1Fu;
                }
                *reinterpret_cast<uint4*>(s_pos + lseg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            } else {
                marks = exact_strip<K, M>(s_packed, wbase + lseg, f.mm_seed, s_pos + lseg);
            }
            // lseg is a multiple of 16: the 32 local positions straddle at most two mask words
            const int sh = lseg & 16;
            atomicOr(&s_minmask[lseg >> 5], marks << sh);
            if (sh && (marks >> 16)) atomicOr(&s_minmask[(lseg >> 5) + 1], marks >> 16);
        }
    }
    __syncwarp();

    // ---------------------------------------------------------------- C: rank the minimizers ----
    // lane l owns mask word l: list index of every marked position
    uint32_t my_word = s_minmask[lane];
    uint32_t n_mine = __popc(my_word);
    uint32_t inc = n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += v;
    }
    const uint32_t n_min = __shfl_sync(0xFFFFFFFFu, inc, 31);
    const uint32_t my_first = inc - n_mine;
    s_wpre[lane] = uint16_t(my_first);

    uint64_t* out = b.codes + a.tile_out[tile];
    const bool chunked = n_min > kCap;
    // does any start of this warp's range yield no code?  (31 mask words, one per lane)
    const bool warp_has_invalid = __any_sync(0xFFFFFFFFu, lane < kWarpKmers / 32 && s_invalid[(wbase >> 5) + lane] != 0);

    // ---------------------------------------------------------------- D + E ----------------------
    for (uint32_t i0 = 0; i0 < n_min; i0 += kCap) {
        const uint32_t i1 = i0 + kCap < n_min ? i0 + kCap : n_min;
        {
            uint32_t word = my_word, idx = my_first;
            while (word) {
                int bit = __ffs(word) - 1;
                word &= word - 1;
                if (idx >= i0 && idx < i1) s_list[idx - i0] = uint16_t(lane * 32 + bit);
                ++idx;
            }
        }
        __syncwarp();
        // kProbes probes per lane in flight: the dependent gathers (pilot rank -> hashed pilot ->
        // [free slot] -> bucket word) of different probes overlap instead of queueing up
        bool special = false, have = false;
        uint32_t hi0 = 0;
        {
            const DevPhf& P = f.minimizer_order;
            const uint32_t nu = (i1 - i0 + 31) >> 5;  // 32-probe groups in this chunk (uniform)
            int bp[kProbes];
            uint64_t h[kProbes];
            uint32_t slot[kProbes];
#pragma unroll
            for (int u = 0; u < kProbes; ++u) {
                if (u < nu) {
                    const uint32_t li = lane + 32 * u;
                    bp[u] = s_list[li < i1 - i0 ? li : 0];  // dead lanes redo entry 0 (harmless)
                    h[u] = murmur64(mmer_at<M>(s_packed, wbase + bp[u]), P.seed);
                    slot[u] = phf_bucket(P, h[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kProbes; ++u)
                if (u < nu) slot[u] = phf_pilot_rank(P, slot[u]);
            uint64_t hp[kProbes];
#pragma unroll
            for (int u = 0; u < kProbes; ++u)
                if (u < nu) hp[u] = __ldg(P.hashed_pilots + slot[u]);
#pragma unroll
            for (int u = 0; u < kProbes; ++u)
                if (u < nu) slot[u] = phf_table_slot(P, h[u] ^ hp[u]);
#pragma unroll
            for (int u = 0; u < kProbes; ++u)
                if (u < nu && slot[u] >= uint32_t(P.num_keys)) slot[u] = __ldg(P.free32 + (slot[u] - uint32_t(P.num_keys)));
            // bucket word -> {B, ns}: hval = base + slope * (bp - q) = (base + slope * bp) + (-slope) * q
            uint64_t word[kProbes];
#pragma unroll
            for (int u = 0; u < kProbes; ++u) {
                if (u < nu) {
                    if (f.buckets.wide) word[u] = __ldg(reinterpret_cast<const uint64_t*>(f.buckets.entries) + slot[u]);
                    else word[u] = __ldg(reinterpret_cast<const uint32_t*>(f.buckets.entries) + slot[u]);
                }
            }
            const int fsh = f.buckets.wide ? 62 : 30;
            const uint64_t bmask = (uint64_t(1) << fsh) - 1;
#pragma unroll
            for (int u = 0; u < kProbes; ++u) {
                const uint32_t li = lane + 32 * u;
                if (u < nu && li < i1 - i0) {
                    const uint32_t flags = uint32_t(word[u] >> fsh);  // bit 1: slope +1, bit 0: colliding
                    const int32_t ns = (flags & 1u) ? 0 : ((flags & 2u) ? -1 : 1);
                    const uint64_t B = (word[u] & bmask) - uint64_t(int64_t(ns) * bp[u]);
                    const uint32_t lo = uint32_t(B), hi = uint32_t(B >> 32);
                    if (u == 0) { hi0 = hi; have = true; }
                    special |= ns == 0 || hi != hi0 || lo - 1024u >= 0xFFFFF800u;  // needs 64-bit care
                    s_ent[li] = Entry{lo, ns};
                    s_hi[li] = hi;
                    s_ref[bp[u]] = uint8_t(li);
                }
            }
        }
        // plain emit needs: every entry of the warp shares lane 0's high word and cannot carry, no
        // colliding minimizer, no start without a code in the warp's range, one chunk
        const uint32_t hi_warp = __shfl_sync(0xFFFFFFFFu, hi0, 0);
        special |= have && hi0 != hi_warp;
        const bool plain = !chunked && !warp_has_invalid && !__any_sync(0xFFFFFFFFu, special);
        __syncwarp();
        if (plain) emit_plain(lane, s_pos, s_ref, s_ent, hi_warp, out + wbase - s_invpre[wbase >> 5]);
        else if (!chunked) emit_general<false>(lane, wbase, i0, i1, s_pos, s_ref, s_minmask, s_wpre, s_ent, s_hi, s_list, s_n_fb, s_invalid, s_invpre, out);
        else emit_general<true>(lane, wbase, i0, i1, s_pos, s_ref, s_minmask, s_wpre, s_ent, s_hi, s_list, s_n_fb, s_invalid, s_invpre, out);
        __syncwarp();

        // colliding minimizers: every k-mer of the run goes through fallback_kmer_order
        // (partitioned_mphf.cpp:308-313, partitioned_mphf.hpp:132-134)
        const uint32_t n_fb = *s_n_fb;
        for (uint32_t e = lane; e < n_fb; e += 32) {
            const int g = wbase + s_list[e];
            const int wi = g >> 4, r = (g & 15) * 2;
            uint32_t x[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) x[j] = (j <= NW) ? s_packed[min(wi + j, kPackedSlots - 1)] : 0u;
            uint32_t y[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) y[j] = __funnelshift_l(x[j + 1], x[j], r);
            // y[0..3] = 128-bit window starting at base g (y[0] most significant)
            uint64_t top = (uint64_t(y[0]) << 32) | y[1], bot = (uint64_t(y[2]) << 32) | y[3];
            uint64_t klo, khi;
            if constexpr (K <= 32) {
                klo = top >> (64 - 2 * K);
                khi = 0;
                (void)bot;
            } else {
                constexpr int sh = 128 - 2 * K;  // 2..62
                klo = (bot >> sh) | (top << (64 - sh));
                khi = top >> sh;
            }
            uint32_t mw = s_invalid[g >> 5];
            int oidx = g - int(s_invpre[g >> 5] + __popc(mw & ((1u << (g & 31)) - 1u)));
            out[oidx] = fallback_code(f, klo, khi);
        }
        __syncwarp();
        if (lane == 0) *s_n_fb = 0;
        __syncwarp();
    }
}

// per tile: contig containing its first in-range position + number of valid k-mer starts before it
__global__ void k_tile_setup(const __grid_constant__ DevBatch b, int64_t pos0, uint32_t n_tiles,
                             uint32_t k, uint32_t* tile_c0, uint64_t* tile_out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int64_t p = pos0 + int64_t(t) * kTile;
    if (p < int64_t(b.first_base)) p = int64_t(b.first_base);
    uint64_t lo = 0, hi = b.n_contigs;  // offsets[lo] <= p (offsets[0] = first_base <= p)
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (int64_t(__ldg(b.offsets + mid)) <= p) lo = mid; else hi = mid;
    }
    uint64_t s = __ldg(b.offsets + lo), e = __ldg(b.offsets + lo + 1);
    uint64_t len = e - s, cnt = len >= k ? len - k + 1 : 0;
    uint64_t before = uint64_t(p) - s;
    tile_c0[t] = uint32_t(lo);
    tile_out[t] = __ldg(b.code_off + lo) + (before < cnt ? before : cnt);
}

template <int K, int M>
void launch_cfg(DevImage const& img, DevBatch const& b, TileArgs const& a, cudaStream_t stream) {
    static bool configured = false;  // per instantiation; the attribute is per device function
    if (!configured) {
        cudaFuncSetAttribute(k_query_tiled<K, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        configured = true;
    }
    k_query_tiled<K, M><<<a.n_tiles, kThreads, kSmemBytes, stream>>>(img, b, a);
}

}  // namespace

uint64_t query_tiled_ws_bytes(uint64_t span_bases) {
    uint64_t n_tiles = (span_bases + 16) / kTile + 2;
    return n_tiles * 12 + 64;
}

bool launch_query_tiled(DevImage const& img, DevBatch const& b, cudaStream_t stream) {
    const uint32_t k = img.k, m = img.m;
    void (*fn)(DevImage const&, DevBatch const&, TileArgs const&, cudaStream_t) = nullptr;
    if (k == 31 && m == 20) fn = launch_cfg<31, 20>;
    else if (k == 31 && m == 16) fn = launch_cfg<31, 16>;
    else if (k == 31 && m == 15) fn = launch_cfg<31, 15>;
    else if (k == 31 && m == 17) fn = launch_cfg<31, 17>;
    else if (k == 31 && m == 18) fn = launch_cfg<31, 18>;
    else if (k == 31 && m == 19) fn = launch_cfg<31, 19>;
    else if (k == 31 && m == 21) fn = launch_cfg<31, 21>;
    else if (k == 21 && m == 11) fn = launch_cfg<21, 11>;
    else if (k == 15 && m == 7) fn = launch_cfg<15, 7>;
    if (!fn || !b.tile_ws || b.n_contigs == 0 || b.n_contigs >= (1ull << 32)) return false;
    if (b.end_base <= b.first_base) return true;
    TileArgs a{};
    const char* p = b.bases + b.first_base;
    uint32_t ali = uint32_t(reinterpret_cast<uintptr_t>(p) & 15u);
    a.abase = p - ali;
    a.pos0 = int64_t(b.first_base) - int64_t(ali);
    uint64_t span = uint64_t(int64_t(b.end_base) - a.pos0);
    a.n_tiles = uint32_t((span + kTile - 1) / kTile);
    if (query_tiled_ws_bytes(b.end_base - b.first_base) > b.tile_ws_bytes) return false;
    uint64_t* tile_out = reinterpret_cast<uint64_t*>(b.tile_ws);
    uint32_t* tile_c0 = reinterpret_cast<uint32_t*>(tile_out + a.n_tiles + 1);
    a.tile_out = tile_out;
    a.tile_c0 = tile_c0;
    k_tile_setup<<<(a.n_tiles + 255) / 256, 256, 0, stream>>>(b, a.pos0, a.n_tiles, k, tile_c0, tile_out);
    fn(img, b, a, stream);
    return true;
}

}  // namespace lphb
