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

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kS = 16;                        // k-mer starts per thread = bases per packed word
constexpr int kLanes = 31;                    // producing lanes per warp
constexpr int kTile = kWarps * kLanes * kS;   // 3968 k-mer starts per tile
constexpr int kMaskWords = kTile / 32;        // 124
constexpr int kPosSlots = kTile + 32;         // positions a minimizer can sit at: kTile + W - 1, rounded
constexpr int kMaskSlots = kPosSlots / 32;    // 125
constexpr int kPackedSlots = 256;
constexpr int kSmemBytes = kPosSlots * (8 + 1 + 1 + 2 + 2) + kPackedSlots * 4 + kMaskSlots * (4 + 4 + 2) + 16;

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

template <int K, int M>
struct Cfg {
    static constexpr int W = K - M + 1;
    static constexpr int NW = (kS + K - 1 + 15) / 16;        // packed words a thread reads
    static constexpr int NH = kS + W - 1;                    // hashes a thread needs
    static constexpr int TileWords = kTile / 16 + NW;        // words staged per tile
    static_assert(W >= 1 && W <= 17, "tiled kernel: window must fit one shuffle hop");
    static_assert(M <= 31 && K <= 63, "k, m out of range");
    static_assert(TileWords <= kThreads, "one load per thread");
};

template <int K, int M>
__global__ void __launch_bounds__(kThreads, 3)
k_query_tiled(const __grid_constant__ DevImage f, const __grid_constant__ DevBatch b,
              const __grid_constant__ TileArgs a) {
    using C = Cfg<K, M>;
    constexpr int W = C::W, NW = C::NW, NH = C::NH;
    static_assert(NH <= 32, "per-thread minimizer marks must fit one 32-bit mask");

    // dynamic shared memory, carved by hand (~58 KB: above the 48 KB static limit)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_base = reinterpret_cast<uint64_t*>(smem_raw);               // per minimizer position: probe result
    int8_t* s_step = reinterpret_cast<int8_t*>(s_base + kPosSlots);         // per minimizer position: +1 / -1 / 0 (collision)
    uint8_t* s_pos = reinterpret_cast<uint8_t*>(s_step + kPosSlots);        // per k-mer: minimizer offset p
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_pos + kPosSlots);      // minimizer positions to probe; later: colliding k-mers
    uint16_t* s_retry = s_list + kPosSlots;                                 // probes that need the free-slot remap
    uint32_t* s_packed = reinterpret_cast<uint32_t*>(s_retry + kPosSlots);  // 2-bit bases
    uint32_t* s_minmask = s_packed + kPackedSlots;     // bit b: position b is some k-mer's minimizer
    uint32_t* s_invalid = s_minmask + kMaskSlots;      // bit q: k-mer start q produces no code
    uint16_t* s_invpre = reinterpret_cast<uint16_t*>(s_invalid + kMaskSlots);  // invalid starts before word
    __shared__ uint32_t s_warp_cnt[kWarps];
    __shared__ uint32_t s_n_min, s_n_retry, s_n_fb, s_n_invalid;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const int64_t T0 = a.pos0 + int64_t(tile) * kTile;  // stream position of tile-local 0
    const int64_t first = int64_t(b.first_base), end = int64_t(b.end_base);

    // ---------------------------------------------------------------- A: load + pack ------------
    if (tid < kMaskSlots) { s_invalid[tid] = 0; s_minmask[tid] = 0; }
    if (tid == 0) { s_n_retry = 0; s_n_fb = 0; s_n_invalid = 0; }
    if (tid >= C::TileWords) s_packed[tid] = 0;
    if (tid < C::TileWords) {
        int64_t wpos = T0 + int64_t(tid) * 16;
        uint32_t word = 0;
        if (wpos + 16 > first && wpos < end) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(a.abase + (wpos - a.pos0)));
            uint32_t y0 = codes4(v.x), y1 = codes4(v.y), y2 = codes4(v.z), y3 = codes4(v.w);
            word = (pack4(y0) << 24) | (pack4(y1) << 16) | (pack4(y2) << 8) | pack4(y3);
            uint32_t bad = bad4(v.x, y0) | bad4(v.y, y1) | bad4(v.z, y2) | bad4(v.w, y3);
            if (bad) {  // rare: flag the contig of every in-range invalid byte
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
        }
        s_packed[tid] = word;
    }
    __syncthreads();

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

    // ---------------------------------------------------------------- B: per-thread scan --------
    const int seg = (warp * kLanes + lane) * kS;  // tile-local position of this thread's first k-mer
    {
        uint32_t wds[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) wds[j] = s_packed[(seg >> 4) + j];

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
            *reinterpret_cast<uint4*>(s_pos + seg) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            // seg is a multiple of 16: the 32 local positions straddle at most two mask words
            const int sh = seg & 16;
            atomicOr(&s_minmask[seg >> 5], marks << sh);
            if (sh && (marks >> 16)) atomicOr(&s_minmask[(seg >> 5) + 1], marks >> 16);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- C: list of minimizers -----
    {
        uint32_t word = tid < kMaskSlots ? s_minmask[tid] : 0u;
        uint32_t n_mine = __popc(word);
        uint32_t inc = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp_cnt[warp] = inc;
        __syncthreads();
        uint32_t warp_base = 0, total = 0;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi) {
            uint32_t v = s_warp_cnt[wi];
            if (wi < warp) warp_base += v;
            total += v;
        }
        uint32_t o = warp_base + inc - n_mine;
        while (word) {
            int bit = __ffs(word) - 1;
            word &= word - 1;
            s_list[o++] = uint16_t(tid * 32 + bit);
        }
        if (tid == 0) s_n_min = total;
    }
    __syncthreads();

    // ---------------------------------------------------------------- D: probe ------------------
    // One thread per distinct minimizer position: m-mer -> PTHash bucket (probes that land in the
    // free-slot region are queued and remapped densely afterwards, single_phf.hpp:61-63) ->
    // wavelet tree -> sizes_and_positions.  Result indexed by minimizer position.
    const uint32_t n_min = s_n_min;
    for (uint32_t idx = tid; idx < n_min; idx += kThreads) {
        const int bp = s_list[idx];
        const int wi = bp >> 4, r = (bp & 15) * 2;
        uint32_t w0 = s_packed[wi], w1 = s_packed[wi + 1], w2 = s_packed[wi + 2];
        uint32_t hi = __funnelshift_l(w1, w0, r), lo = __funnelshift_l(w2, w1, r);
        uint64_t mm = ((uint64_t(hi) << 32) | lo) >> (64 - 2 * M);
        uint64_t pos = phf_raw_position(f.minimizer_order, murmur64(mm, f.minimizer_order.seed));
        if (pos < f.minimizer_order.num_keys) {
            Probe pr = probe_bucket(f, pos);
            s_base[bp] = pr.base;
            s_step[bp] = int8_t(pr.slope);
        } else {
            s_base[bp] = pos;
            s_retry[atomicAdd(&s_n_retry, 1u)] = uint16_t(bp);
        }
    }
    __syncthreads();
    {
        const uint32_t n_retry = s_n_retry;
        for (uint32_t idx = tid; idx < n_retry; idx += kThreads) {
            const int bp = s_retry[idx];
            uint64_t pos = ef_access(f.minimizer_order.free_slots, s_base[bp] - f.minimizer_order.num_keys);
            Probe pr = probe_bucket(f, pos);
            s_base[bp] = pr.base;
            s_step[bp] = int8_t(pr.slope);
        }
        if (n_retry) __syncthreads();
    }

    // ---------------------------------------------------------------- E: codes ------------------
    // k-mer q with minimizer offset p: code = base(q + p) +/- p (partitioned_mphf.cpp:297-337);
    // position-parallel, coalesced 8-byte stores straight to the output.
    uint64_t* out = b.codes + a.tile_out[tile];
    const bool all_valid = s_n_invalid == 0;
    if (all_valid) {
        // 15.5 rounds of 256 consecutive k-mers: fully unrolled, 32-bit shared-memory indexing
        uint64_t* o = out + tid;
        constexpr int kRounds = (kTile + kThreads - 1) / kThreads;
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int q = tid + r * kThreads;
            if (r * kThreads + kThreads <= kTile || q < kTile) {
                const uint32_t p = s_pos[q];
                const int bp = q + int(p);
                const int32_t st = s_step[bp];
                const int32_t sp = st > 0 ? int32_t(p) : -int32_t(p);
                const uint64_t code = s_base[bp] + uint64_t(int64_t(sp));
                if (st != 0) __stcs(o + r * kThreads, code);
                else s_list[atomicAdd(&s_n_fb, 1u)] = uint16_t(q);  // colliding minimizer: needs the k-mer
            }
        }
    } else {
        for (int q = tid; q < kTile; q += kThreads) {
            const uint32_t mw = s_invalid[q >> 5];
            if ((mw >> (q & 31)) & 1u) continue;
            const uint32_t p = s_pos[q];
            const int bp = q + int(p);
            const int32_t st = s_step[bp];
            const int32_t sp = st > 0 ? int32_t(p) : -int32_t(p);
            const uint64_t code = s_base[bp] + uint64_t(int64_t(sp));
            const int oidx = q - int(s_invpre[q >> 5] + __popc(mw & ((1u << (q & 31)) - 1u)));
            if (st != 0) __stcs(out + oidx, code);
            else s_list[atomicAdd(&s_n_fb, 1u)] = uint16_t(q);
        }
    }
    __syncthreads();

    // colliding minimizers: every k-mer of the run goes through fallback_kmer_order
    // (partitioned_mphf.cpp:308-313, partitioned_mphf.hpp:132-134)
    {
        const uint32_t n_fb = s_n_fb;
        for (uint32_t e = tid; e < n_fb; e += kThreads) {
            const int q = s_list[e];
            const int wi = q >> 4, r = (q & 15) * 2;
            uint32_t x[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) x[j] = (j <= NW) ? s_packed[min(wi + j, kPackedSlots - 1)] : 0u;
            uint32_t y[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) y[j] = __funnelshift_l(x[j + 1], x[j], r);
            // y[0..3] = 128-bit window starting at base q (y[0] most significant)
            uint64_t top = (uint64_t(y[0]) << 32) | y[1], bot = (uint64_t(y[2]) << 32) | y[3];
            uint64_t klo, khi;
            if (K <= 32) {
                klo = top >> (64 - 2 * K);
                khi = 0;
            } else {
                constexpr int sh = 128 - 2 * K;  // 2..62
                klo = (bot >> sh) | (top << (64 - sh));
                khi = top >> sh;
            }
            int oidx = q;
            if (!all_valid) {
                uint32_t mw = s_invalid[q >> 5];
                oidx = q - int(s_invpre[q >> 5] + __popc(mw & ((1u << (q & 31)) - 1u)));
            }
            out[oidx] = f.collision_base + fallback_order(f, klo, khi);
        }
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
