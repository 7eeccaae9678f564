// Streaming-query kernel (the hot path) and its build-side (scan) form, for the (k, m) pairs listed
// in pick_cfg(): window widths W = k-m+1 up to 49.
//
// The concatenated base stream is cut into WARP TILES of 2 strips x (32-E) lanes x 16 k-mer starts,
// E = number of lanes to the right a thread's windows reach into (1 for W <= 17: 992 starts per tile;
// 2 for W <= 33: 960; 3 for W <= 49: 928).  The grid is persistent (as many CTAs as fit the GPU);
// every WARP walks its own sequence of tiles (tile = global warp id, + number of warps, ...) and owns
// all the shared memory it touches, so there is no CTA barrier anywhere: a warp stalled on a probe
// never holds up the others.  Per tile:
//
//  A  stage+pack   the tile's ASCII bytes (incl. k-1 bases of overlap) arrive in shared memory by ONE
//                  TMA bulk copy (cp.async.bulk -> mbarrier), issued one tile ahead (double buffer),
//                  so the global-load latency is never exposed.  Each lane then packs 16-byte words
//                  to 32-bit words of 2-bit codes (first base in the most significant bits, the
//                  reference's m-mer / k-mer orientation, partitioned_mphf.hpp:106-108).  Non-ACGT
//                  bytes flag their contig dirty (it is then recomputed by the exact sequential
//                  kernel).  Contig seams are rasterised into a bitmask of k-mer starts
//                  that produce no code (skipped entirely when the tile lies inside one contig, which
//                  the set-up kernel records).
//  B  scan         each thread owns 16 consecutive k-mer starts.  16 m-mer hashes from registers
//                  (MurmurHash2-64, seeded); each is reduced to a 32-bit key = top bits of the hash |
//                  thread-local position (5 bits, 6 for wide windows), so that ONE unsigned min picks
//                  the smaller hash and, between equal key prefixes, the leftmost.
//                  W <= 17: the W-1 keys a thread lacks come from lane+1 by warp shuffle; sliding
//                  minimum = sparse table of 3-input minima (VIMNMX3).  Wider windows: prefix / suffix
//                  minima of the thread's own keys + one neighbour prefix value (and whole-block
//                  minima) per window by shuffle, position field rebased by 16 per lane.
//                  A second pass with the position bits complemented finds the RIGHTMOST minimum; if
//                  the two differ anywhere two candidates shared a key prefix (true repeat or
//                  truncation tie) and that thread recomputes its 16 windows from the full 64-bit
//                  hashes (out of line, rare).  Result: minimizer offset of every k-mer + mask of
//                  positions that are some k-mer's minimizer.
//                  [scan form: the offsets are written out, one byte per valid start, and the tile ends]
//  C  compact      minimizer positions of the tile -> dense list in shared memory (warp scan).
//  D  probe        one lane per distinct minimizer position, kProbes in flight per lane:
//                  m-mer -> PTHash (one 8-byte gather) -> bucket table (one 4/8-byte gather)
//                  (device_mphf.cuh) -> {B, ns} with  code(k-mer at q) = B + ns * q.
//  E  emit         position-parallel: lane l handles k-mers l, l+32, ...: two byte loads find the
//                  entry, one multiply-add makes the code, 8-byte stores are fully coalesced.  Three
//                  forms: plain (every start has a code), masked (contig seams: the invalid-start mask
//                  predicates the store and compacts the index), general (64-bit care, colliding
//                  minimizers -> fallback_kmer_order, partitioned_mphf.cpp:308-313; several chunks).
//
// Every k-mer's code is a pure function of its own k bases, so tiles only share
// k-1 bases of read overlap and nothing else.
//
// Code size matters here: the I-cache holds 32 KB per SM and the warps of an SM sit in different
// phases, so everything rare (tie re-runs, invalid-start rasterisation, general emit, fallback
// k-mers, dirty marking) is __noinline__ and the D loop keeps only kProbes = 2 probes per lane in
// flight (profiles/r01c_experiments.md).
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "device_mphf.cuh"
#include "query_kernels.cuh"

namespace lphb {

uint64_t query_tiled_ws_bytes(uint64_t span_bases);

namespace {

constexpr int kWarps = 4;                      // warps per CTA (independent of each other)
constexpr int kThreads = kWarps * 32;
constexpr int kS = 16;                         // k-mer starts per thread = bases per packed word
constexpr int kStrips = 2;                     // passes per tile
// Per (k, m) (Cfg below): E = lanes to the right a thread's windows reach into (1 for W <= 17, up to 3
// for W <= 49); the last E lanes of a warp only feed their neighbours, so a strip has (32 - E) * 16
// k-mer starts and a tile 2 strips: 992 (E = 1), 960 (E = 2) or 928 (E = 3) starts.
constexpr int kMinTile = kStrips * 29 * kS;    // smallest tile (bounds the set-up workspace)
constexpr int kSlots = 1024;                   // positions a minimizer of the tile's k-mers can sit at (tile + < 64)
constexpr int kSlotWords = kSlots / 32;        // 32: one mask word per lane
// tuned on B200 (profiles/r02_experiments.md): 4 CTAs of 4 warps per SM (128 registers), 3 probes per lane in flight
#ifndef LPHB_PROBES
#define LPHB_PROBES 3
#endif
#ifndef LPHB_MINB
#define LPHB_MINB 4
#endif
#ifndef LPHB_MINB_WIDE
#define LPHB_MINB_WIDE 4
#endif
#ifndef LPHB_EMIT_UNROLL
#define LPHB_EMIT_UNROLL 8
#endif
constexpr int kProbes = LPHB_PROBES;           // probes a lane keeps in flight
constexpr int kEmitUnroll = LPHB_EMIT_UNROLL;  // rows of the plain emit in flight
#ifndef LPHB_LISTCAP
#define LPHB_LISTCAP 256
#endif
constexpr int kListCap = LPHB_LISTCAP;         // minimizer positions probed per D round-trip (chunk); tests build 32
// What a probe leaves for the emit, indexed by the tile-local POSITION of the minimizer:
// code of the k-mer at tile-local q = lo + nsq * q (32-bit; the high word is 0).  nsq = 0 marks an
// entry the fast emit cannot use: lo = 0 colliding minimizer (every k-mer -> fallback_kmer_order),
// lo = 1 the code does not fit 32 bits or may wrap (recomputed exactly in 64 bits, rare).
struct Entry {
    uint32_t lo;
    int32_t nsq;                               // -1: LEFT/MAXIMAL, +1: RIGHT/NONE, 0: see above
};
constexpr int kSpecCap = 60;                   // irregular entries per tile the fix-up pass takes (more: whole tile the slow way)
constexpr int kPackedSlots = 68;               // packed words per tile: 64 + zero read-ahead of mmer_at / win16
constexpr int kRawBytes = 1024;                // ASCII bytes of one tile incl. overlap (64 words), one TMA copy
// per-warp shared memory (bytes)
constexpr int kOffEnt = 0;                               // Entry[kSlots], by minimizer position
constexpr int kOffPos = kOffEnt + kSlots * 8;            // u8[kSlots]  per k-mer start: thread-local minimizer position
constexpr int kOffList = kOffPos + kSlots;               // u16[kListCap] minimizer positions of the chunk
constexpr int kOffMin = kOffList + kListCap * 2;         // u32[32] minimizer-position mask
constexpr int kOffFb = kOffMin + 128;                    // u32 count + u16[kSpecCap]: irregular entries for the fix-up pass
constexpr int kOffInv = kOffFb + 128;                    // u32[32] invalid starts
constexpr int kOffInvPre = kOffInv + 128;                // u16[32] invalid starts before mask word
constexpr int kOffPacked = kOffInvPre + 64;              // u32[kPackedSlots]
constexpr int kOffRaw = (kOffPacked + kPackedSlots * 4 + 15) / 16 * 16;  // kRawBytes (TMA destination)
constexpr int kOffBar = kOffRaw + kRawBytes;             // mbarrier
constexpr int kWarpBytes = kOffBar + 16;
constexpr int kSmemBytes = kWarps * kWarpBytes;
static_assert(kSlotWords == 32 && kWarpBytes % 16 == 0 && kOffRaw % 16 == 0 && kRawBytes % 16 == 0, "per-warp layout");

// what the set-up kernel records per tile
struct __align__(16) TileRec {
    uint64_t out;      // number of valid k-mer starts (codes) before the tile
    uint32_t c0;       // contig containing the tile's first in-range position
    uint32_t clean;    // 1: every start of the tile yields a code (one contig covers tile + k-1 bases)
};

struct TileArgs {
    const char* abase;          // 16-byte aligned; stream position pos0 lives here
    int64_t pos0;               // stream position of abase (may be < first_base by < 16)
    const TileRec* recs;        // per tile
    uint32_t n_tiles;
    // build-side (scan) form only
    uint8_t* records;           // packed 18-byte mm_record_t, scan order
    uint32_t* start_pos;        // per record: stream position of its first k-mer, relative to first_base
    uint64_t* desc;             // 2 words per tile (aggregate, inclusive prefix), zeroed before the launch
    const uint64_t* id_base;    // per contig: m-mer ordinal of its first m-mer
    unsigned long long* n_records;  // total number of records (written by the last tile)
};

// 4 ASCII bytes -> 2-bit codes in the low bits of each byte.  (x>>1 ^ x>>2) & 3 maps
// A,a->0 C,c->1 G,g->2 T,t,U,u->3 (src/constants.cpp:5-13 for the valid bytes).
__device__ __forceinline__ uint32_t codes4(uint32_t x) { return ((x >> 1) ^ (x >> 2)) & 0x03030303u; }
// codes of 4 bytes -> 8 bits, byte 0 in the top 2 bits, left in the TOP byte of the product
__device__ __forceinline__ uint32_t pack4_top(uint32_t y) { return y * 0x40100401u; }
// nonzero iff one of the 4 bytes is not in {A,C,G,T,a,c,g,t}: rebuild the canonical upper-case letter
// of each code with a byte permute and compare.  U/u also come out nonzero here: the (out-of-line,
// exact) slow check that follows accepts them, so RNA input is only slower, not different.
__device__ __forceinline__ uint32_t bad4(uint32_t x, uint32_t y) {
    uint32_t z = y | (y >> 4);
    uint32_t sel = __byte_perm(z, 0u, 0x4420u);       // nibble i = code of byte i
    uint32_t c1 = __byte_perm(0x54474341u, 0u, sel);  // "ACGT"[code]
    return (x & 0xDFDFDFDFu) ^ c1;
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

// ---- TMA bulk copy (global -> shared) completing on an mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the base stream is read once: evict-first in L2, so that it does not displace the image
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t l2_stream_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

template <int K, int M>
struct Cfg {
    static constexpr int W = K - M + 1;
    static constexpr int NW = (kS + K - 1 + 15) / 16;        // packed words a thread reads
    static constexpr int NH = kS + W - 1;                    // m-mers under a thread's 16 windows
    static constexpr int E = (kS - 1 + W - 1) / 16 < 1 ? 1 : (kS - 1 + W - 1) / 16;  // neighbour lanes a window reaches
    static constexpr int Lanes = 32 - E;                     // producing lanes per warp
    static constexpr int Strip = Lanes * kS;                 // k-mer starts per warp pass
    static constexpr int Tile = kStrips * Strip;             // k-mer starts per (warp) tile
    static constexpr int MaskWords = Tile / 32;              // words of the invalid-start mask
    static constexpr int PosBits = E == 1 ? 5 : 6;           // thread-local minimizer position field of a key
    static constexpr int TileWords = (Tile + K - 1 + 15) / 16;  // 16-byte words staged per tile
    static_assert(W >= 1 && E <= 3 && Tile >= kMinTile, "tiled kernel: window must fit three shuffle hops");
    static_assert(NH <= (1 << PosBits), "thread-local minimizer positions must fit the key's position field");
    static_assert(Tile % 32 == 0 && Tile + (E == 1 ? 32 : 64) <= kSlots, "mask words: one per lane");
    static_assert(M <= 31 && K <= 63, "k, m out of range");
    static_assert(TileWords <= 64 && TileWords * 16 <= kRawBytes, "tile must fit its buffers");
    static_assert(Tile / 16 + NW + 2 <= kPackedSlots, "packed read-ahead must stay inside the zero tail");
};

// (hi:lo) * 0xc6a4a7935bd1e995 mod 2^64 in three multiply-adds (IMAD.WIDE + 2 IMAD, all on the FMA pipe)
__device__ __forceinline__ void mul_murmur(uint32_t lo, uint32_t hi, uint32_t& rlo, uint32_t& rhi) {
    asm("{\n"
        ".reg .u64 w;\n"
        ".reg .u32 t;\n"
        "mul.wide.u32 w, %2, 0x5bd1e995;\n"
        "mov.b64 {%0, t}, w;\n"
        "mad.lo.u32 t, %2, 0xc6a4a793, t;\n"
        "mad.lo.u32 %1, %3, 0x5bd1e995, t;\n"
        "}"
        : "=r"(rlo), "=r"(rhi)
        : "r"(lo), "r"(hi));
}
// high word only of the same product (IMAD.HI + 2 IMAD)
__device__ __forceinline__ uint32_t mul_murmur_hi(uint32_t lo, uint32_t hi) {
    return __umulhi(lo, 0x5bd1e995u) + lo * 0xc6a4a793u + hi * 0x5bd1e995u;
}
// top 32 bits of MurmurHash2-64(v, seed) (device_mphf.cuh: murmur64; h0 = seed ^ 8*M): the final
// h ^= h >> 47 cannot change them, and x ^= x >> 47 only touches the low word (x >> 47 has 17 bits)
__device__ __forceinline__ uint32_t murmur_top32(uint32_t v_lo, uint32_t v_hi, uint32_t h0_lo, uint32_t h0_hi) {
    uint32_t xl, xh;
    mul_murmur(v_lo, v_hi, xl, xh);
    xl ^= xh >> 15;
    mul_murmur(xl, xh, xl, xh);
    xl ^= h0_lo;
    xh ^= h0_hi;
    mul_murmur(xl, xh, xl, xh);
    xl ^= xh >> 15;
    return mul_murmur_hi(xl, xh);
}

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

// k-mer starting at tile-local base g as {lo, hi} words of kmer_t (first base most significant,
// include/partitioned_mphf.hpp:108-109)
template <int K, int NW>
__device__ __forceinline__ void kmer_at(const uint32_t* s_packed, int g, uint64_t& klo, uint64_t& khi) {
    const int wi = g >> 4, sh2 = (g & 15) * 2;
    uint32_t x[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) x[j] = (j <= NW) ? s_packed[min(wi + j, kPackedSlots - 1)] : 0u;
    uint32_t y[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) y[j] = __funnelshift_l(x[j + 1], x[j], sh2);
    // y[0..3] = 128-bit window starting at base g (y[0] most significant)
    const uint64_t top = (uint64_t(y[0]) << 32) | y[1], bot = (uint64_t(y[2]) << 32) | y[3];
    if constexpr (K <= 32) {
        klo = top >> (64 - 2 * K);
        khi = 0;
        (void)bot;
    } else {
        constexpr int sh = 128 - 2 * K;  // 2..62
        klo = (bot >> sh) | (top << (64 - sh));
        khi = top >> sh;
    }
}

// Exact minimizer offsets of the 16 k-mers starting at tile-local base g0, from the full 64-bit
// hashes (strict '<' keeps the leftmost on ties: partitioned_mphf.hpp:124,152,159).  Taken only by
// threads whose 27-bit keys tied; out of line.  Returns the mask of minimizer positions.
template <int K, int M>
static __device__ __noinline__ uint64_t exact_strip(const uint32_t* s_packed, int g0, uint64_t seed,
                                                    uint8_t* pos_out) {
    constexpr int W = K - M + 1, NH = kS + W - 1;
    uint64_t h[NH];
#pragma unroll 1
    for (int j = 0; j < NH; ++j) h[j] = murmur64(mmer_at<M>(s_packed, g0 + j), seed);
    uint64_t marks = 0;
#pragma unroll 1
    for (int i = 0; i < kS; ++i) {
        int best = i;
        for (int j = i + 1; j < i + W; ++j)
            if (h[j] < h[best]) best = j;
        pos_out[i] = uint8_t(best);
        marks |= uint64_t(1) << best;
    }
    return marks;
}

// Codes of the k-mers of a tile, position-parallel: lane l handles k-mers l, l+32, ...; one byte
// load gives the position of the k-mer's minimizer, one 8-byte load its entry (the table is indexed
// by position), code = lo + nsq * q, coalesced 8-byte streaming stores.
//
// Plain form, for a tile whose starts all yield a code and whose entries are all regular (32-bit,
// no colliding minimizer): one 32-bit multiply-add per code.
template <int kTile>
__device__ __forceinline__ void emit_plain(int lane, const uint8_t* s_pos, const Entry* s_ent, uint64_t* out_tile) {
    const uint8_t* pos_l = s_pos + lane;
    const unsigned char* ent_l = reinterpret_cast<const unsigned char*>(s_ent + (lane & 16));
    uint2* o = reinterpret_cast<uint2*>(out_tile + lane);
#pragma unroll kEmitUnroll
    for (int r = 0; r < kTile / 32; ++r) {
        // s_pos holds the minimizer's position relative to the owning thread's first start (q & ~15)
        const int2 e = *reinterpret_cast<const int2*>(ent_l + (uint32_t(pos_l[r * 32]) * 8u + r * 256));
        uint32_t lo;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(lo) : "r"(e.y), "r"(lane + r * 32), "r"(e.x));
        __stcs(o + r * 32, make_uint2(lo, 0u));
    }
}

// Masked form: like the plain form, for a tile that does contain starts without a code (contig
// seams: every tile of a batch of short reads): the (warp-uniform) word of the invalid-start mask
// predicates the store and compacts the output index.
template <int kTile>
__device__ __forceinline__ void emit_masked(int lane, const uint8_t* s_pos, const Entry* s_ent,
                                            const uint32_t* s_invalid, const uint16_t* s_invpre, uint64_t* out_tile) {
    const uint8_t* pos_l = s_pos + lane;
    const unsigned char* ent_l = reinterpret_cast<const unsigned char*>(s_ent + (lane & 16));
    const uint32_t lt = (1u << lane) - 1u;
    uint2* o = reinterpret_cast<uint2*>(out_tile + lane);
#pragma unroll 2
    for (int r = 0; r < kTile / 32; ++r) {
        const uint32_t mw = s_invalid[r];  // uniform in the warp
        const int2 e = *reinterpret_cast<const int2*>(ent_l + (uint32_t(pos_l[r * 32]) * 8u + r * 256));
        uint32_t lo;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(lo) : "r"(e.y), "r"(lane + r * 32), "r"(e.x));
        if (!((mw >> lane) & 1u)) __stcs(o + (r * 32 - int(s_invpre[r]) - __popc(mw & lt)), make_uint2(lo, 0u));
    }
}

// k-mer starts of a tile that produce no code (contig seams, short contigs, positions outside
// [first, end)) -> s_invalid.  Only for tiles that one contig does not cover: out of line, away from
// the hot loop's instruction-cache footprint.
template <int K, int kTile>
static __device__ __noinline__ void mark_invalid(DevBatch const& b, uint32_t* s_invalid, int lane, int64_t T0,
                                                 uint32_t c0, int n_starts = kTile) {
    const int kMaskWords = (n_starts + 31) / 32;  // (the scan form also asks about the start after the tile)
    const int64_t first = int64_t(b.first_base), end = int64_t(b.end_base);
    const int64_t tile_end = T0 + n_starts;
    if (T0 < first) {  // head padding of the first tile (< 16 positions)
        int n = int(first - T0);
        if (lane == 0) atomicOr(&s_invalid[0], (1u << n) - 1u);
    }
    uint64_t c = c0;
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
        if (lane >= (lo >> 5) && lane < kMaskWords) {
            uint32_t bits = 0xFFFFFFFFu;
            if (lane == (lo >> 5)) bits <<= (lo & 31);
            atomicOr(&s_invalid[lane], bits);
        }
    }
    __syncwarp();
}

// Exact code of the k-mer starting at tile-local g whose minimizer sits at tile-local mp, without the
// 32-bit shortcuts: the probe in full, hval mod 2^64 like the reference (partitioned_mphf.cpp:337;
// non-members may underflow, SURVEY.md E1 vii), colliding minimizers through fallback_kmer_order
// (partitioned_mphf.cpp:308-313, partitioned_mphf.hpp:132-134).  Used where the fast emit cannot be.
template <int K, int M>
static __device__ __noinline__ uint64_t exact_code(DevImage const& f, const uint32_t* s_packed, int g, int mp) {
    const Probe pr = probe_minimizer(f, mmer_at<M>(s_packed, mp));
    if (pr.slope != 0) return probe_hval(pr, uint32_t(mp - g));
    uint64_t klo, khi;
    kmer_at<K, Cfg<K, M>::NW>(s_packed, g, klo, khi);
    return f.collision_base + fallback_order(f, klo, khi);
}

// where the code of valid start q goes (dense order: starts without a code are skipped)
__device__ __forceinline__ int out_index(const uint32_t* s_invalid, const uint16_t* s_invpre, int q) {
    const uint32_t mw = s_invalid[q >> 5];
    return q - int(s_invpre[q >> 5]) - __popc(mw & ((1u << (q & 31)) - 1u));
}

// Fix-up pass after the fast emit: the k-mers of the (few) minimizers whose entry is not regular
// (s_spec: their positions).  A minimizer at position bp can only serve the starts bp-W+1 .. bp: one
// lane per candidate start, two minimizers per pass when the window fits half a warp.  Out of line.
template <int K, int M, int kTile>
static __device__ __noinline__ void fix_special(DevImage const& f, const uint32_t* s_packed, const uint8_t* s_pos,
                                                const uint16_t* s_spec, uint32_t n_spec, const uint32_t* s_invalid,
                                                const uint16_t* s_invpre, int lane, uint64_t* out) {
    constexpr int W = K - M + 1;
    constexpr int G = W <= 16 ? 16 : 32;   // lanes per minimizer
    constexpr int R = (W + G - 1) / G;     // rounds of G candidate starts
#pragma unroll 1
    for (uint32_t s0 = 0; s0 < n_spec; s0 += 32 / G) {
        const uint32_t s = s0 + uint32_t(lane / G);
        const int bp = s < n_spec ? int(s_spec[s]) : -1;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const int j = (lane % G) + r * G;
            const int q = bp - j;
            if (bp < 0 || j >= W || q < 0 || q >= kTile) continue;
            if ((s_invalid[q >> 5] >> (q & 31)) & 1u) continue;
            if (int(s_pos[q]) + (q & ~15) != bp) continue;  // q's minimizer is another position
            out[out_index(s_invalid, s_invpre, q)] = exact_code<K, M>(f, s_packed, q, bp);
        }
    }
    __syncwarp();
}

// A whole tile the slow way, one lane per start and a full probe per k-mer: tiles with more irregular
// entries than the fix-up pass takes.  Out of line.
template <int K, int M, int kTile>
static __device__ __noinline__ void slow_tile(DevImage const& f, const uint32_t* s_packed, const uint8_t* s_pos,
                                              const uint32_t* s_invalid, const uint16_t* s_invpre, int lane,
                                              uint64_t* out) {
#pragma unroll 1
    for (int q = lane; q < kTile; q += 32) {
        if ((s_invalid[q >> 5] >> (q & 31)) & 1u) continue;
        out[out_index(s_invalid, s_invpre, q)] = exact_code<K, M>(f, s_packed, q, int(s_pos[q]) + (q & ~15));
    }
    __syncwarp();
}

// ---- build-side form: records straight from the tile ----------------------------------------------
// (minimizer::from_string, include/minimizer.hpp:11-170, as the stateless definition SURVEY.md S2':
// one record per maximal run of consecutive k-mers of a contig with the same minimizer occurrence)
// descriptor words are self-contained 64-bit values (nothing else is read on their authority), so
// relaxed device-scope accesses are enough: no fence on either side
__device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr uint64_t kDescValid = uint64_t(1) << 63;
constexpr int kRecChunk = 256;  // records staged in shared memory at a time

// first start >= x that yields a k-mer (bit clear in s_invalid); the caller guarantees there is one
// at or before the tile's last start
__device__ __forceinline__ int next_valid_start(const uint32_t* s_invalid, int x) {
    int w = x >> 5;
    uint32_t bits = ~s_invalid[w] & (0xFFFFFFFFu << (x & 31));
    while (!bits) bits = ~s_invalid[++w];
    return (w << 5) + __ffs(bits) - 1;
}

// One tile of the build-side scan after phase B (s_pos: minimizer position of every start).  A record
// belongs to the tile in which its run ENDS; a run that is still open at the tile's last start is
// finished by the next tile, which learns where it began from this tile's descriptor.  Record
// indices come from a decoupled look-back over the per-tile record counts (chained scan: every tile
// publishes its count at once and the inclusive prefix as soon as its predecessors' are known).
template <int K, int M>
__device__ __forceinline__ void scan_tail(DevBatch const& b, TileArgs const& a, uint64_t seed, uint32_t tile,
                                          int64_t T0, TileRec const& cur, bool tile_clean, int lane,
                                          const uint8_t* s_pos, const uint32_t* s_invalid, const uint32_t* s_packed,
                                          unsigned char* scratch) {
    using C = Cfg<K, M>;
    constexpr int kTile = C::Tile, W = C::W;
    constexpr int kRows = (kTile + 31) / 32;  // mask words holding starts 0 .. kTile-1
    uint16_t* s_ends = reinterpret_cast<uint16_t*>(scratch);              // u16[kSlots]: last start of every run that ends here
    uint16_t* s_rec = reinterpret_cast<uint16_t*>(scratch + 2 * kSlots);  // kRecChunk records of 9 u16
    // V: starts that yield a k-mer (lane r holds word r; bit kTile = the start after the tile)
    const uint32_t inv_w = s_invalid[lane];
    uint32_t V = ~inv_w;
    if (lane * 32 > kTile) V = 0;
    else if (lane * 32 + 31 > kTile) V &= (2u << (kTile - lane * 32)) - 1u;
    // D: bit q set iff start q+1 has another minimizer than start q (absolute position = s_pos + q & ~15)
    uint32_t D = 0;
#pragma unroll 4
    for (int r = 0; r < kRows; ++r) {
        const int q = lane + r * 32;
        const int p0 = int(s_pos[q]) + (q & ~15), p1 = int(s_pos[q + 1]) + ((q + 1) & ~15);
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, p0 != p1);
        if (lane == r) D = bal;
    }
    // does the run of the tile's last start go on into the next tile?  The start after the tile keeps
    // minimizer P of the last start iff P is still inside its window and the m-mer that enters on the
    // right is not smaller (strict '<' moves the minimum: include/minimizer.hpp:94)
    bool cont = false;
    {
        const bool both = !((s_invalid[(kTile - 1) >> 5] >> ((kTile - 1) & 31)) & 1u) && !((s_invalid[kTile >> 5] >> (kTile & 31)) & 1u);
        if (both) {
            const int P = int(s_pos[kTile - 1]) + ((kTile - 1) & ~15);
            if (P >= kTile && lane == 0)
                cont = murmur64(mmer_at<M>(s_packed, kTile + W - 1), seed) >= murmur64(mmer_at<M>(s_packed, P), seed);
            cont = __shfl_sync(0xFFFFFFFFu, cont ? 1 : 0, 0) != 0;
        }
    }
    if (lane == (kTile - 1) / 32) D = (D & ~(1u << ((kTile - 1) & 31))) | ((cont ? 0u : 1u) << ((kTile - 1) & 31));
    // E: run ends = valid starts whose successor is invalid or has another minimizer
    const uint32_t V_up = __shfl_down_sync(0xFFFFFFFFu, V, 1);            // word r+1
    const uint32_t Vnext = (V >> 1) | ((lane < 31 ? V_up : 0u) << 31);     // bit q = V(q + 1)
    uint32_t E = V & (~Vnext | D);
    if (lane * 32 >= kTile) E = 0;
    else if (lane * 32 + 32 > kTile) E &= (1u << (kTile - lane * 32)) - 1u;
    const uint32_t n_mine = __popc(E);
    uint32_t inc = n_mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += v;
    }
    const uint32_t count = __shfl_sync(0xFFFFFFFFu, inc, 31);
    {
        uint16_t* lp = s_ends + (inc - n_mine);
        uint32_t word = E;
        while (word) {  // ascending: the list is in scan order
            const int bit = __ffs(word) - 1;
            word &= word - 1;
            *lp++ = uint16_t(lane * 32 + bit);
        }
    }
    __syncwarp();
    // descriptor: [63] valid, [28] a run is open at the tile's end, [16..27] where it began, [0..15] count
    uint32_t carry_out = 0;
    if (cont) carry_out = uint32_t(next_valid_start(s_invalid, count ? int(s_ends[count - 1]) + 1 : 0));
    uint64_t* descA = a.desc + 2 * uint64_t(tile);
    if (lane == 0) st_desc(descA, kDescValid | (uint64_t(cont ? 1 : 0) << 28) | (uint64_t(carry_out) << 16) | count);
    // contig of the tile (a clean tile lies inside one contig; otherwise every record looks its own up)
    uint64_t c_start = 0, c_idb = 0;
    if (tile_clean) {
        c_start = __ldg(b.offsets + cur.c0);
        c_idb = __ldg(a.id_base + cur.c0);
    }
    const bool v0 = !(s_invalid[0] & 1u);
    int16_t* s_qs = reinterpret_cast<int16_t*>(s_rec + kRecChunk * 9);  // per staged record: its first start (tile-local)
    uint64_t excl = 0;
    bool open_in = false, looked = false;
    int carry_in = 0;
    // The look-back waits for the predecessors, so everything that does not need its result goes
    // first: the records of a chunk are built in shared memory (the first one provisionally: whether it
    // began in the previous tile is the predecessor's news), then the prefix is fetched, then they leave.
#pragma unroll 1
    for (uint32_t e0 = 0; e0 < count || !looked; e0 += kRecChunk) {
        const uint32_t e1 = e0 + kRecChunk < count ? e0 + kRecChunk : count;
#pragma unroll 1
        for (uint32_t e = e0 + lane; e < e1; e += 32) {
            const int q_e = s_ends[e];
            int q_s;  // tile-local start of the run's first k-mer
            if (e == 0) q_s = v0 ? 0 : next_valid_start(s_invalid, 0);
            else q_s = next_valid_start(s_invalid, int(s_ends[e - 1]) + 1);
            const int P = int(s_pos[q_e]) + (q_e & ~15);
            const uint64_t mm = mmer_at<M>(s_packed, P);
            uint64_t cs = c_start, idb = c_idb;
            if (!tile_clean) {
                const uint64_t at = uint64_t(T0 + q_e);
                uint64_t lo = cur.c0, hi = b.n_contigs;  // offsets[lo] <= at < offsets[hi]
                while (hi - lo > 1) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (__ldg(b.offsets + mid) <= at) lo = mid; else hi = mid;
                }
                cs = __ldg(b.offsets + lo);
                idb = __ldg(a.id_base + lo);
            }
            const uint64_t id = idb + (uint64_t(T0 + P) - cs);
            uint16_t* q = s_rec + (e - e0) * 9;  // 18-byte packed mm_record_t (include/constants.hpp:26-33)
            q[0] = uint16_t(mm); q[1] = uint16_t(mm >> 16); q[2] = uint16_t(mm >> 32); q[3] = uint16_t(mm >> 48);
            q[4] = uint16_t(id); q[5] = uint16_t(id >> 16); q[6] = uint16_t(id >> 32); q[7] = uint16_t(id >> 48);
            q[8] = uint16_t(uint32_t(P - q_s) | (uint32_t(q_e - q_s + 1) << 8));  // p1, size
            s_qs[e - e0] = int16_t(q_s);
        }
        if (!looked) {
            looked = true;
            // look back: exclusive prefix of the counts; the predecessor also says whether its last run is open
            if (tile > 0) {
                int64_t j0 = int64_t(tile) - 1;
                bool first_window = true;
                for (;;) {
                    const int64_t j = j0 - lane;
                    uint64_t av = 0, pv = 0;
                    if (j >= 0) {
                        do av = ld_desc(a.desc + 2 * j); while (!(av & kDescValid));
                        pv = ld_desc(a.desc + 2 * j + 1);
                    }
                    if (first_window) {
                        const uint64_t a1 = __shfl_sync(0xFFFFFFFFu, av, 0);
                        open_in = (a1 >> 28) & 1u;
                        carry_in = int((a1 >> 16) & 0xFFFu);
                        first_window = false;
                    }
                    const uint32_t have_p = __ballot_sync(0xFFFFFFFFu, j >= 0 && (pv & kDescValid));
                    const int stop = have_p ? __ffs(have_p) - 1 : 32;  // nearest predecessor whose inclusive prefix is known
                    uint64_t contrib = 0;
                    if (j >= 0 && lane < stop) contrib = av & 0xFFFFu;
                    else if (lane == stop) contrib = pv & ~kDescValid;
#pragma unroll
                    for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, o);
                    excl += contrib;
                    if (have_p || j0 < 32) break;
                    j0 -= 32;
                }
            }
            if (lane == 0) {
                st_desc(descA + 1, kDescValid | (excl + count));
                if (tile + 1 == a.n_tiles) *a.n_records = excl + count;
            }
            __syncwarp();
            if (count && v0 && open_in && lane == 0) {  // the first run began in the previous tile
                const int q_s = carry_in - kTile, q_e = s_ends[0];
                const int P = int(s_pos[q_e]) + (q_e & ~15);
                s_rec[8] = uint16_t(uint32_t(P - q_s) | (uint32_t(q_e - q_s + 1) << 8));
                s_qs[0] = int16_t(q_s);
            }
        }
        __syncwarp();
        if (e1 > e0) {
            for (uint32_t e = e0 + lane; e < e1; e += 32)
                a.start_pos[excl + e] = uint32_t(uint64_t(T0 + int(s_qs[e - e0])) - b.first_base);
            uint16_t* dst = reinterpret_cast<uint16_t*>(a.records) + (excl + e0) * 9;  // records are 2-byte aligned
            const uint32_t n16 = (e1 - e0) * 9;
            for (uint32_t t = lane; t < n16; t += 32) dst[t] = s_rec[t];
        }
        __syncwarp();
    }
}

// kScan = true: the build-side form (minimizer::from_string, include/minimizer.hpp:11-170).  Phases A
// and B, then scan_tail: the packed 18-byte records of the super-k-mers that end in the tile, written in
// scan order at the index a chained scan over the tiles assigns (no per-k-mer array ever reaches
// global memory).  `f` then only carries k, m and the seed; the grid is one warp per tile (not
// persistent), so that tiles start in index order and the look-back stays short.
template <int K, int M, bool kScan = false>
__global__ void __launch_bounds__(kThreads, (Cfg<K, M>::E == 1 ? LPHB_MINB : LPHB_MINB_WIDE))
k_query_tiled(const __grid_constant__ DevImage f, const __grid_constant__ DevBatch b,
              const __grid_constant__ TileArgs a) {
    using C = Cfg<K, M>;
    constexpr int W = C::W, NW = C::NW, NH = C::NH;
    constexpr int kTile = C::Tile, kStrip = C::Strip, kLanes = C::Lanes, kMaskWords = C::MaskWords;

    // dynamic shared memory: one private region per warp, tile-local coordinates
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem_raw + warp * kWarpBytes;
    Entry* s_ent = reinterpret_cast<Entry*>(mine + kOffEnt);          // per minimizer position
    uint8_t* s_pos = mine + kOffPos;                                  // per k-mer: thread-local minimizer position
    uint16_t* s_list = reinterpret_cast<uint16_t*>(mine + kOffList);  // minimizer positions of the chunk
    uint32_t* s_minmask = reinterpret_cast<uint32_t*>(mine + kOffMin);  // bit p: position p is some k-mer's minimizer
    uint32_t* s_spec32 = reinterpret_cast<uint32_t*>(mine + kOffFb);    // [0]: irregular entries of the tile, then their positions (u16)
    uint16_t* s_spec = reinterpret_cast<uint16_t*>(s_spec32 + 1);
    uint32_t* s_invalid = reinterpret_cast<uint32_t*>(mine + kOffInv);  // bit q: k-mer start q produces no code
    uint16_t* s_invpre = reinterpret_cast<uint16_t*>(mine + kOffInvPre);  // invalid starts before mask word
    uint32_t* s_packed = reinterpret_cast<uint32_t*>(mine + kOffPacked);  // 2-bit bases of the tile
    unsigned char* s_raw = mine + kOffRaw;                            // raw ASCII of the tile (TMA destination)
    uint64_t* s_mbar = reinterpret_cast<uint64_t*>(mine + kOffBar);

    const int64_t end = int64_t(b.end_base);
    const uint32_t n_warps = gridDim.x * kWarps;
    uint32_t tile = blockIdx.x * kWarps + warp;
    // bytes of tile t that exist (whole 16-byte words up to the one holding the last base)
    auto tile_bytes = [&](uint32_t t) -> uint32_t {
        const int64_t t0 = a.pos0 + int64_t(t) * kTile;
        int64_t words = (end - t0 + 15) >> 4;
        if (words > C::TileWords) words = C::TileWords;
        return uint32_t(words) * 16u;
    };
    const uint64_t stream_pol = l2_stream_policy();
    if (lane == 0) {
        mbar_init(s_mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile < a.n_tiles) {  // first tile of this warp
            const uint32_t nb = tile_bytes(tile);
            mbar_expect_tx(s_mbar, nb);
            tma_load_1d(s_raw, a.abase + int64_t(tile) * kTile, nb, s_mbar, stream_pol);
        }
    }
    if (lane < kPackedSlots - 64) s_packed[64 + lane] = 0;  // read-ahead tail, never written again
    __syncwarp();
    TileRec rec{};
    if (tile < a.n_tiles) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.recs + tile));
        rec.out = (uint64_t(v.y) << 32) | v.x;
        rec.c0 = v.z;
        rec.clean = v.w;
    }
    const uint64_t h0 = f.mm_seed ^ (8 * kMurmurM);
    const uint32_t h0_lo = uint32_t(h0), h0_hi = uint32_t(h0 >> 32);
    uint32_t keymask;  // ~position mask held in a register so that (hash & mask) | position is one LOP3
    asm volatile("mov.u32 %0, %1;" : "=r"(keymask) : "n"(~((1u << C::PosBits) - 1u)));

    uint32_t iter = 0;
#pragma unroll 1
    for (; tile < a.n_tiles; tile += n_warps, ++iter) {
        const int64_t T0 = a.pos0 + int64_t(tile) * kTile;  // stream position of tile-local 0
        const TileRec cur = rec;

        // ------------------------------------------------------------ A: stage + pack ------------
        s_minmask[lane] = 0;
        s_invalid[lane] = 0;
        mbar_wait(s_mbar, iter & 1u);
        const int n_words = int(tile_bytes(tile) >> 4);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int t = lane + 32 * it;
            uint32_t word = 0;
            if (t < n_words) {
                const uint4 v = *reinterpret_cast<const uint4*>(s_raw + t * 16);
                const uint32_t y0 = codes4(v.x), y1 = codes4(v.y), y2 = codes4(v.z), y3 = codes4(v.w);
                // first base in the most significant bits (partitioned_mphf.hpp:106-108)
                word = __byte_perm(__byte_perm(pack4_top(y1), pack4_top(y0), 0x7300),   // . . y1 y0 -> bytes 2,3
                                   __byte_perm(pack4_top(y3), pack4_top(y2), 0x0073), 0x3254);
                const uint32_t bad = bad4(v.x, y0) | bad4(v.y, y1) | bad4(v.z, y2) | bad4(v.w, y3);
                if (bad) mark_dirty(b, v, T0 + int64_t(t) * 16);  // rare
            }
            s_packed[t] = word;
        }
        __syncwarp();
        // the raw bytes are consumed: fetch this warp's next tile into the same buffer (it lands while
        // this tile is scanned and probed) and its set-up record into registers
        const uint32_t next = tile + n_warps;
        if (next < a.n_tiles) {
            if (lane == 0) {
                const uint32_t nb = tile_bytes(next);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(s_mbar, nb);
                tma_load_1d(s_raw, a.abase + int64_t(next) * kTile, nb, s_mbar, stream_pol);
            }
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.recs + next));
            rec.out = (uint64_t(v.y) << 32) | v.x;
            rec.c0 = v.z;
            rec.clean = v.w;
        }

        // k-mer starts that produce no code (contig seams, short contigs, positions outside
        // [first, end)) -> s_invalid; nothing to do when one contig covers the tile and k-1 more bases
        const bool tile_clean = cur.clean != 0;
        if (!tile_clean) mark_invalid<K, kTile>(b, s_invalid, lane, T0, cur.c0, kScan ? kTile + 1 : kTile);
        // exclusive prefix of invalid counts per mask word (one word per lane)
        bool tile_has_invalid = false;
        {
            const uint32_t mw = tile_clean ? 0u : s_invalid[lane];
            tile_has_invalid = __any_sync(0xFFFFFFFFu, mw != 0 && lane < kMaskWords);
            uint32_t cnt = __popc(mw), inc = cnt;
            if (tile_has_invalid) {
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += v;
                }
            }
            s_invpre[lane] = uint16_t(inc - cnt);
        }

        // ------------------------------------------------------------ B: per-thread scan --------
#pragma unroll 1
        for (int strip = 0; strip < kStrips; ++strip) {
            const int lseg = strip * kStrip + lane * kS;  // tile-local position of this thread's first k-mer
            uint32_t wds[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) wds[j] = s_packed[(lseg >> 4) + j];

            // keys: top bits of the m-mer's hash | thread-local position (PosBits low bits)
            constexpr uint32_t kPosMask = (1u << C::PosBits) - 1u;
            uint32_t key[C::E == 1 ? NH : kS];
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
                const uint32_t top = murmur_top32(v_lo, v_hi, h0_lo, h0_hi);
                asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(key[j]) : "r"(top), "r"(keymask), "r"(uint32_t(j)));  // (h & mask) | j
            }

            uint32_t mn[kS];         // per k-mer: key of its minimizer (leftmost among equal key prefixes)
            uint32_t agree = kPosMask;  // stays kPosMask <=> no window had two candidates with one prefix
            if constexpr (C::E == 1) {
                // W <= 17: the W-1 keys a thread lacks come from lane+1 (positions 16..); sliding
                // minimum by sparse table
#pragma unroll
                for (int j = 0; j < W - 1; ++j) key[kS + j] = __shfl_down_sync(0xFFFFFFFFu, key[j], 1) | 16u;
                window_min<W, NH>(key, mn);
                // same windows with the position bits complemented: the minimum is now the RIGHTMOST
                // one among equal prefixes; both agree on every window <=> no two candidates tied
                uint32_t rkey[NH], rmn[kS];
#pragma unroll
                for (int j = 0; j < NH; ++j) rkey[j] = key[j] ^ kPosMask;
                window_min<W, NH>(rkey, rmn);
#pragma unroll
                for (int i = 0; i < kS; ++i) agree &= mn[i] ^ rmn[i];
            } else {
                // wide windows (up to 3 lanes to the right): a window = suffix of my 16 keys, whole
                // 16-key blocks of the next lanes, prefix of one more lane's block.  Prefix/suffix
                // minima are thread-local; a neighbour's value arrives by shuffle and its position
                // field is rebased by 16 per lane.  r* = same with complemented positions (rightmost).
                uint32_t P[kS], S[kS], rP[kS], rS[kS];
                P[0] = key[0];
                rP[0] = key[0] ^ kPosMask;
#pragma unroll
                for (int j = 1; j < kS; ++j) {
                    P[j] = min(P[j - 1], key[j]);
                    rP[j] = min(rP[j - 1], key[j] ^ kPosMask);
                }
                S[kS - 1] = key[kS - 1];
                rS[kS - 1] = key[kS - 1] ^ kPosMask;
#pragma unroll
                for (int j = kS - 2; j >= 0; --j) {
                    S[j] = min(key[j], S[j + 1]);
                    rS[j] = min(key[j] ^ kPosMask, rS[j + 1]);
                }
                uint32_t Bm[C::E], rBm[C::E];  // whole blocks of lanes +1 .. +E-1
#pragma unroll
                for (int d = 1; d < C::E; ++d) {
                    Bm[d] = __shfl_down_sync(0xFFFFFFFFu, P[kS - 1], d) + 16u * d;
                    rBm[d] = __shfl_down_sync(0xFFFFFFFFu, rP[kS - 1], d) - 16u * d;
                }
#pragma unroll
                for (int i = 0; i < kS; ++i) {
                    const int wend = i + W - 1, e = wend >> 4, c = wend & 15;  // compile-time after unrolling
                    uint32_t acc = min(S[i], __shfl_down_sync(0xFFFFFFFFu, P[c], e) + 16u * e);
                    uint32_t racc = min(rS[i], __shfl_down_sync(0xFFFFFFFFu, rP[c], e) - 16u * e);
#pragma unroll
                    for (int d = 1; d < C::E; ++d) {
                        if (d < e) {
                            acc = min(acc, Bm[d]);
                            racc = min(racc, rBm[d]);
                        }
                    }
                    mn[i] = acc;
                    agree &= acc ^ racc;
                }
            }
            if (lane < kLanes) {  // the last E lanes only feed their neighbours
                uint64_t marks = 0;  // bit j: thread-local position j is the minimizer of one of my k-mers
                if (agree == kPosMask) {
                    uint32_t pk[4];
                    // starts without a code (contig seams) do not ask for their minimizer: in a batch of
                    // short reads that is a fifth fewer probes
                    if constexpr (C::E == 1) {
                        const uint32_t inv16 = tile_clean ? 0u : (s_invalid[lseg >> 5] >> (lseg & 16)) & 0xFFFFu;
                        uint32_t m32 = 0;
                        if (inv16 == 0) {
#pragma unroll
                            for (int i = 0; i < kS; ++i) m32 |= __funnelshift_l(0u, 1u, mn[i]);  // 1 << (mn & 31)
                        } else {
#pragma unroll
                            for (int i = 0; i < kS; ++i)
                                if (!((inv16 >> i) & 1u)) m32 |= __funnelshift_l(0u, 1u, mn[i]);
                        }
                        marks = m32;
                    } else {  // (the wide-window kernel is at its code-size limit: every start marks)
#pragma unroll
                        for (int i = 0; i < kS; ++i) marks |= uint64_t(1) << (mn[i] & kPosMask);
                    }
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        uint32_t t0 = __byte_perm(mn[4 * g4], mn[4 * g4 + 1], 0x0040);
                        uint32_t t1 = __byte_perm(mn[4 * g4 + 2], mn[4 * g4 + 3], 0x0040);
                        pk[g4] = __byte_perm(t0, t1, 0x5410) & (kPosMask * 0x01010101u);
                    }
                    *reinterpret_cast<uint4*>(s_pos + lseg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                } else {
                    marks = exact_strip<K, M>(s_packed, lseg, f.mm_seed, s_pos + lseg);
                }
                // lseg is a multiple of 16: the (up to 64) local positions straddle up to three mask words
                const int sh = lseg & 16;
                const uint32_t m0 = uint32_t(marks << sh), m1 = uint32_t(marks >> (32 - sh));
                atomicOr(&s_minmask[lseg >> 5], m0);
                if (m1) atomicOr(&s_minmask[(lseg >> 5) + 1], m1);
                if constexpr (C::E > 1) {
                    const uint32_t m2 = sh ? uint32_t(marks >> 48) : 0u;
                    if (m2) atomicOr(&s_minmask[(lseg >> 5) + 2], m2);
                }
            }
        }
        __syncwarp();

        if constexpr (kScan) {
            scan_tail<K, M>(b, a, f.mm_seed, tile, T0, cur, tile_clean, lane, s_pos, s_invalid, s_packed,
                            reinterpret_cast<unsigned char*>(s_ent));
            __syncwarp();
            continue;
        }

        // ------------------------------------------------------------ C: list the minimizers ----
        // lane l owns mask word l; list slot of its first marked position = marked positions before it
        uint32_t my_word = s_minmask[lane];
        const uint32_t n_mine = __popc(my_word);
        uint32_t inc = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += v;
        }
        const uint32_t n_min = __shfl_sync(0xFFFFFFFFu, inc, 31);
        uint32_t my_next = inc - n_mine;  // list index of my next marked position

        // ------------------------------------------------------------ D: probe ---------------------
        // one lane per distinct minimizer position, in chunks of kListCap (one chunk unless the window
        // is tiny or the hashes of the tile descend); the entries go to s_ent[position]
        if (lane == 0) s_spec32[0] = 0;
        const DevPhf& P = f.minimizer_order;
        const uint64_t keep = l2_keep_policy();
        const bool wide = f.buckets.wide != 0;
#pragma unroll 1
        for (uint32_t i0 = 0; i0 < n_min; i0 += kListCap) {
            const uint32_t i1 = i0 + kListCap < n_min ? i0 + kListCap : n_min;
            while (my_word && my_next < i1) {  // my marked positions that fall into this chunk
                const int bit = 31 - __clz(my_word);  // (any order: the entry table is indexed by position)
                my_word ^= 1u << bit;
                s_list[my_next - i0] = uint16_t(lane * 32 + bit);
                ++my_next;
            }
            __syncwarp();
            const uint32_t n_chunk = i1 - i0;
            const uint32_t nu = (n_chunk + 31) >> 5;  // 32-probe groups in this chunk (uniform)
            // kProbes probes per lane in flight: the dependent gathers (hashed pilot -> bucket word)
            // of different probes overlap instead of queueing up
#pragma unroll 1
            for (uint32_t g0 = 0; g0 < nu; g0 += kProbes) {
                const uint32_t ng = nu - g0;  // live groups of this round (uniform), >= 1
                int bp[kProbes];
                uint64_t h[kProbes];
                uint32_t slot[kProbes];
#pragma unroll
                for (int u = 0; u < kProbes; ++u) {
                    if (u < ng) {
                        const uint32_t li = lane + 32 * (g0 + u);
                        bp[u] = s_list[li < n_chunk ? li : 0];  // dead lanes redo entry 0 (harmless)
                        h[u] = murmur64(mmer_at<M>(s_packed, bp[u]), P.seed);
                        slot[u] = phf_bucket(P, h[u]);
                    }
                }
                uint64_t hp[kProbes];
#pragma unroll
                for (int u = 0; u < kProbes; ++u)
#ifdef LPHB_EXP_NOGATHER  // timing experiment only (wrong codes): no image reads
                    if (u < ng) hp[u] = h[u] + slot[u];
#else
#ifdef LPHB_EXP_SMALLTAB  // timing experiment only (wrong codes): gathers confined to an L2-resident corner
                    if (u < ng) hp[u] = ldg_keep(P.pilot_hash + (slot[u] & 0x3FFFFu), keep);
#else
                    if (u < ng) hp[u] = ldg_keep(P.pilot_hash + slot[u], keep);
#endif
#endif
                // the bucket table is indexed by the raw table slot (free slots folded in at load time)
                uint32_t wlo[kProbes], whi[kProbes];
#pragma unroll
                for (int u = 0; u < kProbes; ++u) {
                    if (u < ng) {
#ifdef LPHB_EXP_SMALLTAB
                        const uint32_t ts = phf_table_slot(P, h[u] ^ hp[u]) & 0xFFFFFu;
#else
                        const uint32_t ts = phf_table_slot(P, h[u] ^ hp[u]);
#endif
#ifdef LPHB_EXP_NOGATHER
                        wlo[u] = (ts & 0x3FFFFFFFu) | 0x80000000u;
                        whi[u] = 0;
                        continue;
#endif
                        if (wide) {
                            const uint64_t wv = ldg_keep(reinterpret_cast<const uint64_t*>(f.buckets.entries) + ts, keep);
                            wlo[u] = uint32_t(wv);
                            whi[u] = uint32_t(wv >> 32);
                        } else {
                            wlo[u] = ldg_keep(reinterpret_cast<const uint32_t*>(f.buckets.entries) + ts, keep);
                            whi[u] = 0;
                        }
                    }
                }
                // bucket word -> entry: hval(q) = base + slope * (bp - q) = (base + slope * bp) - slope * q
#pragma unroll
                for (int u = 0; u < kProbes; ++u) {
                    const uint32_t li = lane + 32 * (g0 + u);
                    if (u < ng && li < n_chunk) {
                        // flags on top of the word: slope +1 (LEFT/MAXIMAL) | colliding; base below
                        const uint32_t top = wide ? whi[u] : wlo[u];
                        const uint32_t base = wide ? wlo[u] : (wlo[u] & 0x3FFFFFFFu);
                        const int32_t nsq = (int32_t(top) >> 31) | 1;  // -slope: -1 if slope +1, else +1
                        const bool colliding = (top & 0x40000000u) != 0;
                        // regular <=> base + slope * o stays inside [0, 2^32) for every offset o <= k - m
                        bool care = nsq > 0 ? base < uint32_t(K - M) : base > 0xFFFFFFFFu - uint32_t(K - M);
                        if (wide) care |= (top & 0x3FFFFFFFu) != 0;
#ifdef LPHB_TEST_CARE_BELOW  // test build: push regular entries through the exact 64-bit path too
                        care |= base < uint32_t(LPHB_TEST_CARE_BELOW);
#endif
                        Entry e;
                        e.lo = base - uint32_t(nsq * bp[u]);
                        e.nsq = nsq;
                        if (colliding | care) {  // rare: the fast emit's code for these k-mers is overwritten afterwards
                            const uint32_t n = atomicAdd(s_spec32, 1u);
                            if (n < uint32_t(kSpecCap)) s_spec[n] = uint16_t(bp[u]);
                        }
                        s_ent[bp[u]] = e;
                    }
                }
            }
            __syncwarp();
        }

        // ------------------------------------------------------------ E: emit -----------------------
        uint64_t* out = b.codes + cur.out;
        if (!tile_has_invalid) emit_plain<kTile>(lane, s_pos, s_ent, out);
        else emit_masked<kTile>(lane, s_pos, s_ent, s_invalid, s_invpre, out);
        // irregular entries (colliding minimizers, codes outside 32 bits): their k-mers again, exactly
        __syncwarp();
        const uint32_t n_spec = s_spec32[0];
        if (n_spec) {
            __syncwarp();
            if (n_spec <= uint32_t(kSpecCap)) fix_special<K, M, kTile>(f, s_packed, s_pos, s_spec, n_spec, s_invalid, s_invpre, lane, out);
            else slow_tile<K, M, kTile>(f, s_packed, s_pos, s_invalid, s_invpre, lane, out);
        }
        __syncwarp();
    }  // tiles of this warp
}

// per tile: contig containing its first in-range position, number of valid k-mer starts before it,
// and whether one contig covers the tile plus k-1 bases (then every start yields a code)
__global__ void k_tile_setup(const __grid_constant__ DevBatch b, int64_t pos0, uint32_t n_tiles,
                             uint32_t k, int kTile, int extra, TileRec* recs) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t t0 = pos0 + int64_t(t) * kTile;
    int64_t p = t0;
    if (p < int64_t(b.first_base)) p = int64_t(b.first_base);
    uint64_t lo = 0, hi = b.n_contigs;  // offsets[lo] <= p (offsets[0] = first_base <= p)
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (int64_t(__ldg(b.offsets + mid)) <= p) lo = mid; else hi = mid;
    }
    uint64_t s = __ldg(b.offsets + lo), e = __ldg(b.offsets + lo + 1);
    uint64_t len = e - s, cnt = len >= k ? len - k + 1 : 0;
    uint64_t before = uint64_t(p) - s;
    TileRec r;
    r.c0 = uint32_t(lo);
    r.out = __ldg(b.code_off + lo) + (before < cnt ? before : cnt);
    r.clean = (t0 >= int64_t(b.first_base) && int64_t(s) <= t0 && int64_t(e) >= t0 + kTile + extra + int64_t(k) - 1) ? 1u : 0u;
    recs[t] = r;
}

// CTAs of one instantiation that fit a device at once, looked up per device (a process may drive
// several GPUs from several host threads)
template <int K, int M, bool kScan>
int resident_ctas() {
    static std::mutex mu;
    static int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (!cached[dev]) {
        cudaFuncSetAttribute(k_query_tiled<K, M, kScan>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_query_tiled<K, M, kScan>, kThreads, kSmemBytes);
        cached[dev] = sms * (per_sm > 0 ? per_sm : 1);
    }
    return cached[dev];
}

template <int K, int M, bool kScan>
void launch_cfg(DevImage const& img, DevBatch const& b, TileArgs const& a, cudaStream_t stream) {
    const int resident = resident_ctas<K, M, kScan>();
    // persistent grid: every warp walks tiles (global warp id) + i * (number of warps)
    const uint32_t ctas_needed = (a.n_tiles + kWarps - 1) / kWarps;
    // (the scan form is not persistent: tiles must start in index order for its chained scan)
    const uint32_t grid = (kScan || ctas_needed < uint32_t(resident)) ? ctas_needed : uint32_t(resident);
    k_query_tiled<K, M, kScan><<<grid, kThreads, kSmemBytes, stream>>>(img, b, a);
}

using LaunchFn = void (*)(DevImage const&, DevBatch const&, TileArgs const&, cudaStream_t);

// the (k, m) pairs the kernel is instantiated for
bool pick_cfg(uint32_t k, uint32_t m, bool scan, LaunchFn& fn, int& tile) {
    fn = nullptr;
#define LPHB_CFG(KK, MM)                                                   \
    if (k == KK && m == MM) {                                              \
        fn = scan ? launch_cfg<KK, MM, true> : launch_cfg<KK, MM, false>;  \
        tile = Cfg<KK, MM>::Tile;                                          \
    }
    LPHB_CFG(31, 20) LPHB_CFG(31, 16) LPHB_CFG(31, 15) LPHB_CFG(31, 17) LPHB_CFG(31, 18) LPHB_CFG(31, 19)
    LPHB_CFG(31, 21) LPHB_CFG(21, 11) LPHB_CFG(15, 7)
    LPHB_CFG(63, 24) LPHB_CFG(47, 20)  // wide windows (W = 40, 28), 128-bit k-mers
#undef LPHB_CFG
    return fn != nullptr;
}

struct ScanOut {
    uint8_t* records;
    uint32_t* start_pos;
    const uint64_t* id_base;
    unsigned long long* n_records;
};

bool launch_tiled(DevImage const& img, DevBatch const& b, const ScanOut* scan, cudaStream_t stream) {
    LaunchFn fn;
    int tile = 0;  // k-mer starts per tile of the instantiation
    if (!pick_cfg(img.k, img.m, scan != nullptr, fn, tile)) return false;
    if (!b.tile_ws || b.n_contigs == 0 || b.n_contigs >= (1ull << 32)) return false;
    if (b.end_base <= b.first_base) return true;
    TileArgs a{};
    const char* p = b.bases + b.first_base;
    uint32_t ali = uint32_t(reinterpret_cast<uintptr_t>(p) & 15u);
    a.abase = p - ali;
    a.pos0 = int64_t(b.first_base) - int64_t(ali);
    uint64_t span = uint64_t(int64_t(b.end_base) - a.pos0);
    uint64_t n_tiles = (span + tile - 1) / tile;
    if (n_tiles >= (1ull << 31)) return false;
    a.n_tiles = uint32_t(n_tiles);
    if (query_tiled_ws_bytes(b.end_base - b.first_base) > b.tile_ws_bytes) return false;
    TileRec* recs = reinterpret_cast<TileRec*>(b.tile_ws);
    a.recs = recs;
    if (scan) {
        a.records = scan->records;
        a.start_pos = scan->start_pos;
        a.id_base = scan->id_base;
        a.n_records = scan->n_records;
        a.desc = reinterpret_cast<uint64_t*>(recs + n_tiles);  // second half of the workspace
        cudaMemsetAsync(a.desc, 0, n_tiles * 16, stream);
    }
    k_tile_setup<<<(a.n_tiles + 255) / 256, 256, 0, stream>>>(b, a.pos0, a.n_tiles, img.k, tile, scan ? 1 : 0, recs);
    fn(img, b, a, stream);
    return true;
}

}  // namespace

uint64_t query_tiled_ws_bytes(uint64_t span_bases) {
    uint64_t n_tiles = (span_bases + 16) / kMinTile + 2;
    return n_tiles * (sizeof(TileRec) + 16) + 64;  // per tile: its set-up record + (scan form) its chained-scan descriptor
}

bool launch_query_tiled(DevImage const& img, DevBatch const& b, cudaStream_t stream) {
    return launch_tiled(img, b, nullptr, stream);
}

bool launch_scan_records_tiled(uint32_t k, uint32_t m, uint64_t seed, DevBatch const& b, const uint64_t* d_id_base,
                               uint8_t* d_records, uint32_t* d_start_pos, unsigned long long* d_n_records,
                               cudaStream_t stream) {
    DevImage img{};  // the scan form reads k, m and the minimizer seed only
    img.k = k;
    img.m = m;
    img.w = k - m + 1;
    img.mm_seed = seed;
    ScanOut out{d_records, d_start_pos, d_id_base, d_n_records};
    return launch_tiled(img, b, &out, stream);
}

bool scan_tiled_available(uint32_t k, uint32_t m) {
    LaunchFn fn;
    int tile = 0;
    return pick_cfg(k, m, true, fn, tile);
}

}  // namespace lphb
