// Generic query kernels: valid for every (k, m) the format allows.
//   k_query_generic  one thread per k-mer, stateless definition: the code of a
//                    k-mer is query(kmer, leftmost-minimum m-mer, its offset).  It is the
//                    always-available path for (k, m) pairs the tiled kernel (query_tiled.cu) is
//                    not instantiated for.
//   k_query_quirk    one thread per contig: exact sequential restatement of the reference's
//                    streaming loop, used ONLY for contigs that contain non-ACGT bytes, where the
//                    reference's output depends on stale state.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "device_mphf.cuh"
#include "query_kernels.cuh"

namespace lphb {

namespace {

struct KmerCount {
    const uint64_t* offsets;
    uint32_t k;
    __host__ __device__ uint64_t operator()(uint64_t c) const {
        uint64_t len = offsets[c + 1] - offsets[c];
        return len >= k ? len - k + 1 : 0;
    }
};

__global__ void k_finish_code_offsets(const uint64_t* offsets, uint64_t n_contigs, uint32_t k,
                                      uint64_t* code_off, unsigned long long* status) {
    // ExclusiveSum wrote code_off[0..n); close the array and publish the total.
    uint64_t total = 0;
    if (n_contigs) {
        uint64_t len = offsets[n_contigs] - offsets[n_contigs - 1];
        total = code_off[n_contigs - 1] + (len >= k ? len - k + 1 : 0);
    }
    code_off[n_contigs] = total;
    status[0] = total;
    status[1] = 0;
}

// index of the contig containing stream position i (offsets[c] <= i < offsets[c+1])
__device__ __forceinline__ uint64_t find_contig(const uint64_t* offsets, uint64_t n, uint64_t i) {
    uint64_t lo = 0, hi = n;  // invariant: offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_query_generic(const __grid_constant__ DevImage f,
                                                       const __grid_constant__ DevBatch b) {
    const uint32_t k = f.k, m = f.m, w = f.w;
    const uint64_t mm_mask = (uint64_t(1) << (2 * m)) - 1;
    uint64_t span = b.end_base - b.first_base;
    for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < span;
         t += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t i = b.first_base + t;
        uint64_t c = find_contig(b.offsets, b.n_contigs, i);
        uint64_t start = __ldg(b.offsets + c), end = __ldg(b.offsets + c + 1);
        if (i + k > end) continue;
        // forward k-mer, first base in the most significant bits (partitioned_mphf.hpp:106-108)
        uint64_t lo = 0, hi = 0;
        bool ok = true;
        for (uint32_t j = 0; j < k; ++j) {
            uint32_t code = nt4(uint8_t(b.bases[i + j]));
            if (code > 3) { ok = false; break; }
            hi = (hi << 2) | (lo >> 62);
            lo = (lo << 2) | code;
        }
        if (!ok) {
            b.dirty[c] = 1;  // plain store: every writer stores the same value
            continue;
        }
        // leftmost minimum over the w m-mers (strict '<' keeps the leftmost on ties)
        uint64_t best_h = 0, best_mm = 0;
        uint32_t best_p = 0;
        for (uint32_t p = 0; p < w; ++p) {
            uint32_t sh = 2 * (k - m - p);
            uint64_t mm = sh == 0 ? lo : (sh < 64 ? ((lo >> sh) | (hi << (64 - sh))) : (hi >> (sh - 64)));
            mm &= mm_mask;
            uint64_t h = murmur64(mm, f.mm_seed);
            if (p == 0 || h < best_h) { best_h = h; best_mm = mm; best_p = p; }
        }
        Probe pr = probe_minimizer(f, best_mm);
        uint64_t hval = pr.slope == 0 ? pr.base + fallback_order(f, lo, hi) : probe_hval(pr, best_p);
        b.codes[__ldg(b.code_off + c) + (i - start)] = hval;
    }
}

// ---------------------------------------------------------------------------------------------
// Sequential restatement of mphf::operator()(contig, len, streaming=true) INCLUDING its
// behaviour on non-ACGT bytes.  ref: include/partitioned_mphf.hpp:73-184.  Only `run` and the
// ring cursor are reset at an invalid byte (:179-183); the ring, the minimum's slot, p1, the
// rolling k-mer and the last probe context stay, so a rescan of stale slots can push a spurious
// code while m <= run < k.
// ---------------------------------------------------------------------------------------------
__global__ void k_query_quirk(const __grid_constant__ DevImage f, const char* bases,
                              const uint64_t* offsets, const uint64_t* list, uint64_t n_list,
                              const uint64_t* out_off, uint64_t* out, uint64_t* counts) {
    uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (j >= n_list) return;
    uint64_t c = list[j];
    const char* s = bases + offsets[c];
    uint64_t len = offsets[c + 1] - offsets[c];
    uint64_t* dst = out + out_off[j];
    uint64_t n_out = 0;
    const uint32_t k = f.k, m = f.m, w = f.w;
    if (len < k) { counts[j] = 0; return; }
    const uint64_t mm_mask = (uint64_t(1) << (2 * m)) - 1;
    const uint32_t top = 2 * k - 64;  // bits of the k-mer living in `hi` when k > 32
    const uint64_t lo_mask = k >= 32 ? ~uint64_t(0) : ((uint64_t(1) << (2 * k)) - 1);
    const uint64_t hi_mask = k > 32 ? ((uint64_t(1) << top) - 1) : 0;
    uint64_t ring_mm[64], ring_h[64];
    for (uint32_t q = 0; q < w; ++q) ring_mm[q] = 0, ring_h[q] = 0;
    uint32_t cursor = 0, min_slot = w, p1 = 0;
    uint64_t mmer = 0, lo = 0, hi = 0, run = 0;
    // last probe context (mm_context_t, partitioned_mphf.hpp:55-60)
    uint64_t g_rank = 0, l_rank = 0;
    int32_t ctx_slope = 1;  // slope of the last probe (0: colliding minimizer); MAXIMAL before any
    for (uint64_t i = 0; i < len; ++i) {
        uint32_t code = nt4(uint8_t(s[i]));
        if (code > 3) { run = 0; cursor = 0; continue; }
        mmer = ((mmer << 2) | code) & mm_mask;
        hi = ((hi << 2) | (lo >> 62)) & hi_mask;
        lo = ((lo << 2) | code) & lo_mask;
        ++run;
        if (run < m) continue;
        int action = 0;  // 0 keep, 1 rescan, 2 newcomer
        if (cursor == min_slot) action = 1;
        ring_mm[cursor] = mmer;
        ring_h[cursor] = murmur64(mmer, f.mm_seed);
        if (run == k) {
            action = 1;
        } else if (run > k && ring_h[min_slot] > ring_h[cursor]) {
            p1 = k - m;
            min_slot = cursor;
            action = 2;
        }
        if (action == 0) {
            if (run >= k) {
                if (ctx_slope == 0) l_rank = fallback_order(f, lo, hi);
                else if (ctx_slope < 0) ++l_rank;  // RIGHT, NONE
                else --l_rank;                     // LEFT, MAXIMAL
                dst[n_out++] = g_rank + l_rank;
            }
        } else {
            if (action == 1) {
                min_slot = (cursor + 1) % w;
                p1 = 0;
                uint32_t step = 1;
                for (uint32_t q = (cursor + 2) % w; q < w; ++q, ++step)
                    if (ring_h[min_slot] > ring_h[q]) { min_slot = q; p1 = step; }
                for (uint32_t q = 0; q <= (cursor + 2) % w; ++q, ++step)
                    if (ring_h[min_slot] > ring_h[q]) { min_slot = q; p1 = step; }
            }
            Probe pr = probe_minimizer(f, ring_mm[min_slot]);
            ctx_slope = pr.slope;
            // split hval back into the reference's (global_rank, local_rank) so that the
            // +-1 continuation above reproduces :131-145 (mod 2^64)
            if (pr.slope == 0) { g_rank = pr.base; l_rank = fallback_order(f, lo, hi); }
            else if (pr.slope > 0) { g_rank = pr.base; l_rank = p1; }
            else { g_rank = pr.base; l_rank = uint64_t(0) - uint64_t(p1); }
            dst[n_out++] = g_rank + l_rank;
        }
        cursor = (cursor + 1) % w;
    }
    counts[j] = n_out;
}

__global__ void k_assemble(uint64_t* dst, const uint64_t* dst_off, const uint64_t* src_a,
                           const uint64_t* src_b, const uint64_t* src_off, const uint8_t* from_b,
                           uint64_t n_contigs) {
    // one warp per contig
    uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    for (uint64_t c = warp; c < n_contigs; c += (uint64_t(gridDim.x) * blockDim.x) >> 5) {
        uint64_t n = dst_off[c + 1] - dst_off[c];
        const uint64_t* src = (from_b[c] ? src_b : src_a) + src_off[c];
        uint64_t* d = dst + dst_off[c];
        for (uint64_t i = lane; i < n; i += 32) d[i] = src[i];
    }
}

__global__ void k_count_dirty(const uint8_t* dirty, uint64_t n, unsigned long long* status) {
    unsigned long long local = 0;
    for (uint64_t c = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; c < n;
         c += uint64_t(gridDim.x) * blockDim.x)
        local += dirty[c] ? 1 : 0;
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(status + 1, local);
}

}  // namespace

uint64_t code_offsets_tmp_bytes(uint64_t n_contigs) {
    size_t bytes = 0;
    KmerCount op{nullptr, 1};
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::TransformInputIterator<uint64_t, KmerCount, cub::CountingInputIterator<uint64_t>> it(cnt, op);
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (uint64_t*)nullptr, n_contigs);
    return bytes + 256;
}

void launch_code_offsets(const uint64_t* d_offsets, uint64_t n_contigs, uint32_t k,
                         uint64_t* d_code_off, unsigned long long* d_status, void* d_tmp,
                         uint64_t tmp_bytes, cudaStream_t stream) {
    if (n_contigs) {
        KmerCount op{d_offsets, k};
        cub::CountingInputIterator<uint64_t> cnt(0);
        cub::TransformInputIterator<uint64_t, KmerCount, cub::CountingInputIterator<uint64_t>> it(cnt, op);
        size_t bytes = tmp_bytes;
        cub::DeviceScan::ExclusiveSum(d_tmp, bytes, it, d_code_off, n_contigs, stream);
    }
    k_finish_code_offsets<<<1, 1, 0, stream>>>(d_offsets, n_contigs, k, d_code_off, d_status);
}

void launch_query_generic(DevImage const& img, DevBatch const& b, cudaStream_t stream) {
    uint64_t span = b.end_base - b.first_base;
    if (span == 0) return;
    uint64_t blocks = (span + 255) / 256;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    k_query_generic<<<unsigned(blocks), 256, 0, stream>>>(img, b);
}

void launch_query_quirk(DevImage const& img, const char* bases, const uint64_t* offsets,
                        const uint64_t* list, uint64_t n_list, const uint64_t* out_off,
                        uint64_t* out, uint64_t* counts, cudaStream_t stream) {
    if (!n_list) return;
    k_query_quirk<<<unsigned((n_list + 63) / 64), 64, 0, stream>>>(img, bases, offsets, list, n_list,
                                                                 out_off, out, counts);
}

void launch_assemble(uint64_t* dst, const uint64_t* dst_off, const uint64_t* src_a,
                     const uint64_t* src_b, const uint64_t* src_off, const uint8_t* from_b,
                     uint64_t n_contigs, cudaStream_t stream) {
    if (!n_contigs) return;
    uint64_t blocks = (n_contigs * 32 + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    k_assemble<<<unsigned(blocks), 256, 0, stream>>>(dst, dst_off, src_a, src_b, src_off, from_b, n_contigs);
}

void launch_count_dirty(const uint8_t* dirty, uint64_t n, unsigned long long* status,
                        cudaStream_t stream) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    k_count_dirty<<<unsigned(blocks), 256, 0, stream>>>(dirty, n, status);
}

}  // namespace lphb
