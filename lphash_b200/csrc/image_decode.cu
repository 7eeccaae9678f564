// See image_decode.cuh.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <vector>

#include "api_internal.h"
#include "device_mphf.cuh"
#include "image_decode.cuh"

namespace lphb {

namespace {

unsigned grid_for(uint64_t n) {
    uint64_t blocks = (n + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    return unsigned(blocks);
}

// error kinds only decoding reveals; the messages are those of the host decode
enum DecodeError : int {
    kPilotRank = 0, kFreeSlotTable, kFreeSlotKeys, kSpIndex, kBaseBits, kRootOnes, kNumErrors
};
const char* const kErrorText[kNumErrors] = {
    "pilot rank outside dictionary", "single_phf: free slot outside the table", "single_phf: free slot outside [0, num_keys)",
    "sizes_and_positions index out of range", "bucket base does not fit 62 bits",
    "wavelet tree: root ones != size of the max/none leaf"};

struct DCompact {  // compact_vector on the device (words 8-byte aligned, one spare zero word behind them)
    const uint64_t* words;
    uint64_t size, width, mask, nwords;
};
// compact_vector::access (compact_vector.hpp:229-234)
__device__ __forceinline__ uint64_t compact_get(DCompact const& c, uint64_t i) {
    if (c.width == 0) return 0;
    const uint64_t pos = i * c.width, wd = pos >> 6, sh = pos & 63;
    uint64_t v = c.words[wd] >> sh;
    if (sh && wd + 1 < c.nwords) v |= c.words[wd + 1] << (64 - sh);
    return v & c.mask;
}

struct DBits {  // bit vector + ones before every word (rs_bit_vector's rank directory, recomputed)
    const uint64_t* words;  // nw + 1 words, the last one zero
    const uint64_t* cum;    // nw + 1 entries
    uint64_t nbits;
};
__device__ __forceinline__ bool bit_at(DBits const& b, uint64_t i) { return (b.words[i >> 6] >> (i & 63)) & 1; }
// rs_bit_vector::rank (include/rs_bit_vector.hpp:32-43): ones in [0, pos)
__device__ __forceinline__ uint64_t rank1(DBits const& b, uint64_t pos) {
    const uint64_t w = pos >> 6, r = pos & 63;
    return b.cum[w] + (r ? uint64_t(__popcll(b.words[w] & ((uint64_t(1) << r) - 1))) : 0);
}

struct MaskedPop {  // popcount of word i with the bits at and beyond nbits cleared; word nw counts 0
    const uint64_t* words;
    uint64_t nbits, nw;
    __host__ __device__ uint64_t operator()(uint64_t i) const {
        if (i >= nw) return 0;
        uint64_t x = words[i];
        if ((i + 1) * 64 > nbits) x &= nbits > i * 64 ? ((uint64_t(1) << (nbits - i * 64)) - 1) : 0;
#ifdef __CUDA_ARCH__
        return uint64_t(__popcll(x));
#else
        return uint64_t(__builtin_popcountll(x));
#endif
    }
};

// pilots: dual<dictionary, dictionary> (encoders.hpp:167-170, 268-271) + default_hash64(pilot, seed) (single_phf.hpp:58)
__global__ void k_pilot_hash(DCompact fr, DCompact fd, DCompact br, DCompact bd, uint64_t n_buckets, uint64_t seed,
                             uint64_t* out, unsigned* err) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b < n_buckets; b += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t pilot;
        if (b < fr.size) {
            const uint64_t r = compact_get(fr, b);
            if (r >= fd.size) { atomicAdd(err + kPilotRank, 1u); continue; }
            pilot = compact_get(fd, r);
        } else {
            const uint64_t r = compact_get(br, b - fr.size);
            if (r >= bd.size) { atomicAdd(err + kPilotRank, 1u); continue; }
            pilot = compact_get(bd, r);
        }
        out[b] = murmur64(pilot, seed);
    }
}

// Elias-Fano access for every i at once: value i = ((position of the i-th set high bit) - i) << l | low[i]
// (include/ef_sequence.hpp:77-81); one thread per word of the high bits, cum = ones before the word
__global__ void k_ef_values(const uint64_t* high, const uint64_t* cum, uint64_t nw, uint64_t nbits, uint64_t positions,
                            DCompact low, uint64_t* vals) {
    for (uint64_t wi = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; wi < nw; wi += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t x = high[wi];
        if ((wi + 1) * 64 > nbits) x &= nbits > wi * 64 ? ((uint64_t(1) << (nbits - wi * 64)) - 1) : 0;
        uint64_t i = cum[wi];
        while (x && i < positions) {
            const uint64_t pos = wi * 64 + uint64_t(__ffsll((long long)x) - 1);
            x &= x - 1;
            vals[i] = ((pos - i) << low.width) | compact_get(low, i);
            ++i;
        }
    }
}

__global__ void k_free32(const uint64_t* vals, uint64_t n, uint64_t table_size, uint32_t* out, unsigned* err) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        if (vals[i] >= table_size) { atomicAdd(err + kFreeSlotTable, 1u); continue; }
        out[i] = uint32_t(vals[i]);
    }
}

struct EntryWriter {
    void* p;
    uint32_t wide;
    __device__ void set(uint64_t i, bool slope_up, bool colliding, uint64_t base) const {
        if (wide) reinterpret_cast<uint64_t*>(p)[i] = (uint64_t(slope_up) << 63) | (uint64_t(colliding) << 62) | base;
        else reinterpret_cast<uint32_t*>(p)[i] = (uint32_t(slope_up) << 31) | (uint32_t(colliding) << 30) | uint32_t(base & 0x3FFFFFFFu);
    }
    __device__ void copy(uint64_t to, uint64_t from) const {
        if (wide) reinterpret_cast<uint64_t*>(p)[to] = reinterpret_cast<uint64_t*>(p)[from];
        else reinterpret_cast<uint32_t*>(p)[to] = reinterpret_cast<uint32_t*>(p)[from];
    }
};

// One word per bucket id: quartet_wtree::rank_of (src/quartet_wtree.cpp:84-99) + the Elias-Fano reads of
// mphf::query (src/partitioned_mphf.cpp:292-339), see device_mphf.cuh for the five cases
__global__ void k_bucket_words(DBits root, DBits left_right, DBits max_none, const uint64_t* sp, uint64_t sp_n, uint64_t D,
                               uint64_t w, uint64_t maxblock, uint64_t rs, uint64_t ns, uint64_t np, uint32_t k_minus_m,
                               EntryWriter e, unsigned long long* max_base, unsigned* err) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b < D; b += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t n1 = rank1(root, b), n0 = b - n1;  // ones / zeros of the root before b
        const bool msb = bit_at(root, b);
        uint64_t base = 0, kind = 0, i0 = 0, i1 = 0, i2 = 0;
        bool bad = false;
        auto S = [&](uint64_t i) -> uint64_t {
            if (i >= sp_n) { bad = true; return 0; }
            return sp[i];
        };
        (void)i0; (void)i1; (void)i2;
        if (!msb) {
            const bool lsb = bit_at(left_right, n0);
            const uint64_t r_right = rank1(left_right, n0), r_left = n0 - r_right;
            if (!lsb) {  // LEFT
                base = S(r_left) + maxblock;
                kind = 1;
            } else {     // RIGHT or collision (size 0)
                const uint64_t v1 = S(rs + r_right), v2 = S(rs + r_right + 1);
                if (v2 == v1) {
                    kind = 0;
                } else {
                    base = v1 + maxblock + k_minus_m;
                    kind = 2;
                }
            }
        } else {
            const bool lsb = bit_at(max_none, n1);
            const uint64_t r_none = rank1(max_none, n1), r_max = n1 - r_none;
            if (!lsb) {  // MAXIMAL
                base = w * r_max;
                kind = 1;
            } else {     // NONE
                base = S(ns + r_none) + maxblock + (S(np + r_none + 1) - S(np + r_none));
                kind = 2;
            }
        }
        if (bad) { atomicAdd(err + kSpIndex, 1u); continue; }
        if (base >> 62) { atomicAdd(err + kBaseBits, 1u); continue; }
        atomicMax(max_base, (unsigned long long)base);
        e.set(b, kind == 1, kind == 0, base);
    }
}

// mphf_alt: hval(k-mer) = sizes[i] + positions.diff(i) - p; size 0 = colliding (src/unpartitioned_mphf.cpp:193-203)
__global__ void k_bucket_words_alt(const uint64_t* positions, const uint64_t* sizes, uint64_t D, EntryWriter e,
                                   unsigned long long* max_base, unsigned* err) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b < D; b += uint64_t(gridDim.x) * blockDim.x) {
        if (sizes[b + 1] == sizes[b]) {
            e.set(b, false, true, 0);
        } else {
            const uint64_t base = sizes[b] + (positions[b + 1] - positions[b]);
            if (base >> 62) { atomicAdd(err + kBaseBits, 1u); continue; }
            atomicMax(max_base, (unsigned long long)base);
            e.set(b, false, false, base);
        }
    }
}

// table slots >= num_keys repeat the word of the key position free_slots sends them to (single_phf.hpp:61-63)
__global__ void k_fold_free(const uint32_t* free32, uint64_t n_free, uint64_t D, EntryWriter e, unsigned* err) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n_free; i += uint64_t(gridDim.x) * blockDim.x) {
        if (free32[i] >= D) { atomicAdd(err + kFreeSlotKeys, 1u); continue; }
        e.copy(D + i, free32[i]);
    }
}

// ---- host side: uploads and sequencing ---------------------------------------------------------------------------
struct Scratch {  // device buffers of one load, freed at the end
    std::vector<void*> bufs;
    ~Scratch() {
        for (void* p : bufs) cudaFree(p);
    }
    template <class T>
    T* alloc(uint64_t count) {
        void* p = nullptr;
        CK(cudaMalloc(&p, count * sizeof(T) + 64));
        bufs.push_back(p);
        return static_cast<T*>(p);
    }
    // the words of a serialized vector, 8-byte aligned, followed by `spare` zero words
    uint64_t* upload(VecRef const& v, uint64_t spare, cudaStream_t s) {
        uint64_t* d = alloc<uint64_t>(v.n + spare);
        if (v.n) CK(cudaMemcpyAsync(d, v.p, v.n * 8, cudaMemcpyHostToDevice, s));
        if (spare) CK(cudaMemsetAsync(d + v.n, 0, spare * 8, s));
        return d;
    }
};

DCompact compact_on_device(Scratch& sc, CompactRef const& c, cudaStream_t s) {
    DCompact d;
    d.words = sc.upload(c.words, 1, s);
    d.size = c.size;
    d.width = c.width;
    d.mask = c.width ? ((c.width == 64 ? 0 : (uint64_t(1) << c.width)) - 1) : 0;
    d.nwords = c.words.n;
    return d;
}

// ones before every word (nw + 1 entries); returns the device array and the total through *total
uint64_t* prefix_ones(Scratch& sc, const uint64_t* d_words, uint64_t nw, uint64_t nbits, void* tmp, size_t tmp_bytes,
                      cudaStream_t s) {
    uint64_t* cum = sc.alloc<uint64_t>(nw + 1);
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::TransformInputIterator<uint64_t, MaskedPop, cub::CountingInputIterator<uint64_t>> it(cnt, MaskedPop{d_words, nbits, nw});
    size_t bytes = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, bytes, it, cum, int64_t(nw + 1), s);
    return cum;
}

size_t scan_tmp_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::TransformInputIterator<uint64_t, MaskedPop, cub::CountingInputIterator<uint64_t>> it(cnt, MaskedPop{nullptr, 0, 0});
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, static_cast<uint64_t*>(nullptr), int64_t(n + 1));
    return bytes;
}

// every value of an Elias-Fano sequence, on the device
uint64_t* ef_on_device(Scratch& sc, EfRef const& r, void* tmp, size_t tmp_bytes, cudaStream_t s) {
    uint64_t* high = sc.upload(r.high, 1, s);
    uint64_t* cum = prefix_ones(sc, high, r.high.n, r.nbits, tmp, tmp_bytes, s);
    uint64_t total = 0;
    CK(cudaMemcpyAsync(&total, cum + r.high.n, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (total < r.positions) throw FormatError("EF: fewer set bits than values");
    DCompact low = compact_on_device(sc, r.low, s);
    uint64_t* vals = sc.alloc<uint64_t>(r.positions + 2);
    k_ef_values<<<grid_for(r.high.n), 256, 0, s>>>(high, cum, r.high.n, r.nbits, r.positions, low, vals);
    return vals;
}

DBits bits_on_device(Scratch& sc, BitsRef const& r, void* tmp, size_t tmp_bytes, cudaStream_t s) {
    DBits d;
    uint64_t* w = sc.upload(r.words, 1, s);
    d.words = w;
    d.cum = prefix_ones(sc, w, r.words.n, r.nbits, tmp, tmp_bytes, s);
    d.nbits = r.nbits;
    return d;
}

void phf_on_device(Scratch& sc, PhfRef const& r, DevPhf const& phf, uint8_t* arena, void* tmp, size_t tmp_bytes, unsigned* err,
                   cudaStream_t s) {
    DCompact fr = compact_on_device(sc, r.front_ranks, s), fd = compact_on_device(sc, r.front_dict, s),
             br = compact_on_device(sc, r.back_ranks, s), bd = compact_on_device(sc, r.back_dict, s);
    auto* pilot_hash = reinterpret_cast<uint64_t*>(arena + uintptr_t(phf.pilot_hash));
    k_pilot_hash<<<grid_for(r.n_buckets), 256, 0, s>>>(fr, fd, br, bd, r.n_buckets, phf.seed, pilot_hash, err);
    uint64_t* free_vals = ef_on_device(sc, r.free_slots, tmp, tmp_bytes, s);
    auto* free32 = reinterpret_cast<uint32_t*>(arena + uintptr_t(phf.free32));
    k_free32<<<grid_for(r.n_free), 256, 0, s>>>(free_vals, r.n_free, phf.table_size, free32, err);
}

}  // namespace

DevImage rebase_image(DevImage d, const void* device_base) {
    auto* base = static_cast<const uint8_t*>(device_base);
    for (DevPhf* p : {&d.minimizer_order, &d.fallback}) {
        p->free32 = reinterpret_cast<const uint32_t*>(base + uintptr_t(p->free32));
        p->pilot_hash = reinterpret_cast<const uint64_t*>(base + uintptr_t(p->pilot_hash));
    }
    d.buckets.entries = base + uintptr_t(d.buckets.entries);
    return d;
}

void decode_phf_on_device(ImagePlan const& P, void* d_arena) {
    cudaStream_t s = nullptr;
    Scratch sc;
    auto* arena = static_cast<uint8_t*>(d_arena);
    CK(cudaMemsetAsync(arena, 0, P.arena_bytes, s));
    unsigned* err = sc.alloc<unsigned>(kNumErrors + 2);
    CK(cudaMemsetAsync(err, 0, (kNumErrors + 2) * sizeof(unsigned), s));
    const size_t tmp_bytes = scan_tmp_bytes(P.minimizer_order.free_slots.high.n + 1);
    void* tmp = sc.alloc<uint8_t>(tmp_bytes);
    phf_on_device(sc, P.minimizer_order, P.img.minimizer_order, arena, tmp, tmp_bytes, err, s);
    unsigned h_err[kNumErrors] = {};
    CK(cudaMemcpyAsync(h_err, err, sizeof h_err, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    for (int i = 0; i < kNumErrors; ++i)
        if (h_err[i]) throw FormatError(kErrorText[i]);
}

void decode_image_on_device(ImagePlan const& P, void* d_arena, uint64_t* collision_base) {
    cudaStream_t s = nullptr;
    Scratch sc;
    auto* arena = static_cast<uint8_t*>(d_arena);
    CK(cudaMemsetAsync(arena, 0, P.arena_bytes, s));
    unsigned* err = sc.alloc<unsigned>(kNumErrors + 2);
    auto* max_base = reinterpret_cast<unsigned long long*>(sc.alloc<uint64_t>(1));
    CK(cudaMemsetAsync(err, 0, (kNumErrors + 2) * sizeof(unsigned), s));
    CK(cudaMemsetAsync(max_base, 0, 8, s));
    uint64_t longest = std::max(P.minimizer_order.free_slots.high.n, P.fallback.free_slots.high.n);
    for (uint64_t n : {P.root.words.n, P.left_right.words.n, P.max_none.words.n, P.sizes_and_positions.high.n, P.positions.high.n,
                       P.sizes.high.n})
        longest = std::max(longest, n);
    const size_t tmp_bytes = scan_tmp_bytes(longest + 1);
    void* tmp = sc.alloc<uint8_t>(tmp_bytes);

    DevImage const& img = P.img;
    phf_on_device(sc, P.minimizer_order, img.minimizer_order, arena, tmp, tmp_bytes, err, s);
    phf_on_device(sc, P.fallback, img.fallback, arena, tmp, tmp_bytes, err, s);
    const uint64_t D = img.distinct_minimizers;
    EntryWriter e{arena + uintptr_t(img.buckets.entries), img.buckets.wide};
    if (P.alt) {
        uint64_t* positions = ef_on_device(sc, P.positions, tmp, tmp_bytes, s);
        uint64_t* sizes = ef_on_device(sc, P.sizes, tmp, tmp_bytes, s);
        k_bucket_words_alt<<<grid_for(D), 256, 0, s>>>(positions, sizes, D, e, max_base, err);
        *collision_base = img.collision_base;
    } else {
        DBits root = bits_on_device(sc, P.root, tmp, tmp_bytes, s), left_right = bits_on_device(sc, P.left_right, tmp, tmp_bytes, s),
              max_none = bits_on_device(sc, P.max_none, tmp, tmp_bytes, s);
        uint64_t root_ones = 0;
        CK(cudaMemcpyAsync(&root_ones, root.cum + P.root.words.n, 8, cudaMemcpyDeviceToHost, s));
        uint64_t* sp = ef_on_device(sc, P.sizes_and_positions, tmp, tmp_bytes, s);  // synchronizes
        if (root_ones != P.max_none.nbits) throw FormatError(kErrorText[kRootOnes]);
        const uint64_t maxblock = uint64_t(img.w) * img.n_maximal;
        uint64_t sp_np = 0;
        CK(cudaMemcpyAsync(&sp_np, sp + img.none_pos_start, 8, cudaMemcpyDeviceToHost, s));
        k_bucket_words<<<grid_for(D), 256, 0, s>>>(root, left_right, max_none, sp, P.sizes_and_positions.positions, D, img.w, maxblock,
                                                   img.right_start, img.none_sizes_start, img.none_pos_start, img.k - img.m, e,
                                                   max_base, err);
        CK(cudaStreamSynchronize(s));
        *collision_base = sp_np + maxblock;  // partitioned_mphf.cpp:308-311
    }
    k_fold_free<<<grid_for(P.minimizer_order.n_free), 256, 0, s>>>(
        reinterpret_cast<const uint32_t*>(arena + uintptr_t(img.minimizer_order.free32)), P.minimizer_order.n_free, D, e, err);
    unsigned h_err[kNumErrors] = {};
    unsigned long long h_max = 0;
    CK(cudaMemcpyAsync(h_err, err, sizeof h_err, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&h_max, max_base, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    for (int i = 0; i < kNumErrors; ++i)
        if (h_err[i]) throw FormatError(kErrorText[i]);
    if (!img.buckets.wide && h_max >= (1ull << 30)) throw FormatError("bucket base beyond the number of k-mers");
}

}  // namespace lphb
