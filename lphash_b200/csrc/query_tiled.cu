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
//  B  scan        each thread owns 16 consecutive k-mer starts: 16 m-mer hashes from registers
//                 (MurmurHash2-64, seeded), the W-1 it lacks from lane+1 by warp shuffle (lane 31
//                 only feeds lane 30: warps overlap by one lane), van Herk / Gil-Werman sliding
//                 minimum with leftmost ties -> minimizer offset p of every k-mer, and a 16-bit
//                 mask of super-k-mer heads (minimizer occurrence differs from the predecessor's).
//  C  compact     heads of the CTA -> dense list in shared memory (warp scan + per-thread loop).
//  D  probe       one thread per head: minimizer m-mer -> PTHash -> wavelet tree -> Elias-Fano
//                 (device_mphf.cuh) -> the head's hash code and the run's slope are left in the
//                 head's own staging slot.
//  E  fill+store  each thread walks its 16 k-mers: code = previous -/+ 1 inside a super-k-mer
//                 (partitioned_mphf.hpp:131-145), reload at heads; colliding runs are queued and
//                 resolved densely through fallback_kmer_order; codes leave through a padded
//                 shared-memory transpose as fully coalesced 8-byte stores.
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
constexpr int kCap = 256;                      // probe results held at once per warp
constexpr int kWarpBytes = kWarpSlots * (1 + 2 + 2) + kCap * 8 + (kWarpMaskWords + 4) * 4;
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
    static constexpr int NH = kS + W - 1;                    // hashes a thread needs
    static constexpr int TileWords = kTile / 16 + NW;        // words staged per tile
    static_assert(W >= 1 && W <= 17, "tiled kernel: window must fit one shuffle hop");
    static_assert(M <= 31 && K <= 63, "k, m out of range");
    static_assert(TileWords <= kPackedSlots, "packed tile must fit its shared-memory array");
};

// Codes of the k-mers [q0, q1) of a warp whose minimizers are list entries [i0, i1) (their probe
// results sit in s_base[0 .. i1-i0)).  `checked` = some starts of the tile produce no code.
template <bool kChecked, bool kChunked>
__device__ __forceinline__ void emit_codes(int lane, int wbase, uint32_t i0, uint32_t i1,
                                           const uint8_t* s_pos, const uint16_t* s_ref,
                                           const uint64_t* s_base, uint16_t* s_list,
                                           uint32_t* s_n_fb, const uint32_t* s_invalid,
                                           const uint16_t* s_invpre, uint64_t* out) {
    uint64_t* o = out + wbase + lane;
#pragma unroll 4
    for (int r = 0; r < kWarpKmers / 32; ++r) {
        const int q = lane + r * 32;
        const uint32_t p = s_pos[q];
        const uint32_t ref = s_ref[q + int(p)];
        const uint32_t idx = ref & 0x3FFFu, kind = ref >> 14;  // kind 1: +p, 2: -p, 0: collision
        if (kChunked && (idx < i0 || idx >= i1)) continue;
        const int32_t sp = kind == 1 ? int32_t(p) : -int32_t(p);
        const uint64_t code = s_base[idx - i0] + uint64_t(int64_t(sp));
        if (!kChecked) {
            if (kind != 0) __stcs(o + r * 32, code);
            else s_list[atomicAdd(s_n_fb, 1u)] = uint16_t(q);  // colliding minimizer: needs the k-mer
        } else {
            const int g = wbase + q;
            const uint32_t mw = s_invalid[g >> 5];
            if ((mw >> (g & 31)) & 1u) continue;
            const int oidx = g - int(s_invpre[g >> 5] + __popc(mw & ((1u << (g & 31)) - 1u)));
            if (kind != 0) __stcs(out + oidx, code);
            else s_list[atomicAdd(s_n_fb, 1u)] = uint16_t(q);
        }
    }
}

template <int K, int M>
__global__ void __launch_bounds__(kThreads, 6)
k_query_tiled(const __grid_constant__ DevImage f, const __grid_constant__ DevBatch b,
              const __grid_constant__ TileArgs a) {
    using C = Cfg<K, M>;
    constexpr int W = C::W, NW = C::NW, NH = C::NH;
    static_assert(NH <= 32, "per-thread minimizer marks must fit one 32-bit mask");

    // dynamic shared memory: one private region per warp (warp-local coordinates) + tile-wide
    // packed bases and invalid-start bitmask
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* mine = smem_raw + warp * kWarpBytes;
    uint64_t* s_base = reinterpret_cast<uint64_t*>(mine);                  // per probe (list index): result
    uint8_t* s_pos = reinterpret_cast<uint8_t*>(s_base + kCap);            // per k-mer: minimizer offset p
    uint16_t* s_ref = reinterpret_cast<uint16_t*>(s_pos + kWarpSlots);     // per position: list index | kind << 14
    uint16_t* s_list = s_ref + kWarpSlots;                                 // minimizer positions (chunk); later colliding k-mers (<= kWarpKmers)
    uint32_t* s_minmask = reinterpret_cast<uint32_t*>(s_list + kWarpSlots);  // bit b: position b is some k-mer's minimizer
    uint32_t* s_n_fb = s_minmask + kWarpMaskWords;                         // colliding k-mers queued
    uint32_t* s_packed = reinterpret_cast<uint32_t*>(smem_raw + kWarps * kWarpBytes);  // 2-bit bases (tile)
    uint32_t* s_invalid = s_packed + kPackedSlots;     // bit q: k-mer start q produces no code (tile)
    uint16_t* s_invpre = reinterpret_cast<uint16_t*>(s_invalid + kMaskSlots);  // invalid starts before word
    __shared__ uint32_t s_n_invalid;

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
        if (lane == 31) s_n_invalid = inc;
    }
    __syncthreads();  // invalid-start mask ready; from here on every warp runs on its own

    // ---------------------------------------------------------------- B: per-thread scan --------
    const int wbase = warp * kWarpKmers;  // tile-local position of this warp's first k-mer
#pragma unroll 1
    for (int strip = 0; strip < kStrips; ++strip) {
        const int lseg = strip * kStrip + lane * kS;  // warp-local position of this thread's first k-mer
        uint32_t wds[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) wds[j] = s_packed[((wbase + lseg) >> 4) + j];

        uint64_t h[NH];
#pragma unroll
        for (int j = 0; j < kS; ++j) {
            // m-mer starting at base j of the thread's window: top 2M bits of the 64-bit window at j
            const int q = j >> 4, r = (j & 15) * 2;
            uint32_t hi = r ? __funnelshift_l(wds[q + 1], wds[q], r) : wds[q];
            uint32_t lo;
            if (q + 2 < NW) lo = r ? __funnelshift_l(wds[q + 2], wds[q + 1], r) : wds[q + 1];
            else lo = wds[q + 1] << r;
            uint64_t win = (uint64_t(hi) << 32) | lo;
            h[j] = murmur64(win >> (64 - 2 * M), f.mm_seed);
        }
#pragma unroll
        for (int j = 0; j < W - 1; ++j) h[kS + j] = __shfl_down_sync(0xFFFFFFFFu, h[j], 1);

        // van Herk / Gil-Werman over blocks of W hashes: window i = [i, i+W-1] is the suffix of
        // its block from i joined with the prefix of the next block up to i+W-1.  Leftmost wins
        // ties (strict comparisons, partitioned_mphf.hpp:124,152,159).  Fully unrolled with
        // compile-time indices: suf/pre live in registers and unused entries vanish.
        uint64_t suf_h[NH], pre_h[NH];
        uint32_t suf_p[NH], pre_p[NH];
#pragma unroll
        for (int j = NH - 1; j >= 0; --j) {
            if (j % W == W - 1 || j == NH - 1) {
                suf_h[j] = h[j];
                suf_p[j] = j;
            } else {
                bool keep = h[j] <= suf_h[j + 1];  // element j is to the left: it wins ties
                suf_h[j] = keep ? h[j] : suf_h[j + 1];
                suf_p[j] = keep ? uint32_t(j) : suf_p[j + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            if (j % W == 0) {
                pre_h[j] = h[j];
                pre_p[j] = j;
            } else {
                bool take = h[j] < pre_h[j - 1];  // element j is to the right: strictly smaller only
                pre_h[j] = take ? h[j] : pre_h[j - 1];
                pre_p[j] = take ? uint32_t(j) : pre_p[j - 1];
            }
        }
        uint32_t marks = 0;  // bit j: thread-local position j is the minimizer of one of my k-mers
        uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < kS; ++i) {
            uint32_t bpos;
            if (i % W == 0) {
                bpos = suf_p[i];
            } else {
                bool take = pre_h[i + W - 1] < suf_h[i];
                bpos = take ? pre_p[i + W - 1] : suf_p[i];
            }
            marks |= 1u << bpos;
            pk[i >> 2] |= (bpos - i) << (8 * (i & 3));
        }
        if (lane < kLanes) {  // lane 31 only feeds hashes to lane 30
            *reinterpret_cast<uint4*>(s_pos + lseg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            // lseg is a multiple of 16: the 32 local positions straddle at most two mask words
            const int sh = lseg & 16;
            atomicOr(&s_minmask[lseg >> 5], marks << sh);
            if (sh && (marks >> 16)) atomicOr(&s_minmask[(lseg >> 5) + 1], marks >> 16);
        }
    }
    __syncwarp();

    // ---------------------------------------------------------------- C: rank the minimizers ----
    // lane l owns mask word l: list index of every marked position -> s_ref
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

    uint64_t* out = b.codes + a.tile_out[tile];
    const bool all_valid = s_n_invalid == 0;

    // ---------------------------------------------------------------- D + E ----------------------
    // One lane per distinct minimizer position: m-mer -> PTHash -> wavelet tree -> sizes_and_positions
    // (device_mphf.cuh); then k-mer q with minimizer offset p gets base(q + p) +/- p
    // (partitioned_mphf.cpp:297-337), position-parallel, coalesced 8-byte stores.  At most kCap
    // probe results are held at once; denser strips (tiny windows) go round the loop again.
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
        for (uint32_t idx = i0 + lane; idx < i1; idx += 32) {
            const int bp = s_list[idx - i0];
            const int g = wbase + bp;  // tile-local base index of the minimizer
            const int wi = g >> 4, r = (g & 15) * 2;
            uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1], w2 = s_packed[wi + 2];
            uint32_t hi = __funnelshift_l(w1, w0, r), lo = __funnelshift_l(w2, w1, r);
            uint64_t mm = ((uint64_t(hi) << 32) | lo) >> (64 - 2 * M);
            Probe pr = probe_minimizer(f, mm);
            s_base[idx - i0] = pr.base;
            s_ref[bp] = uint16_t(idx | ((pr.slope > 0 ? 1u : (pr.slope < 0 ? 2u : 0u)) << 14));
        }
        __syncwarp();
        const bool chunked = n_min > kCap;
        if (!chunked) {
            if (all_valid) emit_codes<false, false>(lane, wbase, i0, i1, s_pos, s_ref, s_base, s_list, s_n_fb, s_invalid, s_invpre, out);
            else emit_codes<true, false>(lane, wbase, i0, i1, s_pos, s_ref, s_base, s_list, s_n_fb, s_invalid, s_invpre, out);
        } else {
            // colliding k-mers are queued in s_list, which the next chunk reuses: flush per chunk
            emit_codes<true, true>(lane, wbase, i0, i1, s_pos, s_ref, s_base, s_list, s_n_fb, s_invalid, s_invpre, out);
        }
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
