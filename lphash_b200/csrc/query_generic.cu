// Generic query kernels: valid for every (k, m) the format allows.
//   k_query_generic  one thread per k-mer, stateless definition: the code of a
//                    k-mer is query(kmer, leftmost-minimum m-mer, its offset).  It is the
//                    always-available path for (k, m) pairs the tiled kernel (query_tiled.cu) is
//                    not instantiated for.
// (contigs with non-ACGT bytes, where the reference's output depends on stale state - SURVEY.md Q1 -
//  are finished by quirk_kernels.cu.)
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

// bytes outside ACGT/acgt/U/u -> 'A' (what the reference's non-streaming branch makes of them:
// include/mphf_utils.hpp:108, seq_nt4_table[c] & 3)
__global__ void k_sanitize(char* bases, uint64_t n) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x)
        if (nt4(uint8_t(bases[i])) > 3) bases[i] = 'A';
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

void launch_sanitize(char* d_bases, uint64_t n, cudaStream_t stream) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    k_sanitize<<<unsigned(blocks), 256, 0, stream>>>(d_bases, n);
}

void launch_count_dirty(const uint8_t* dirty, uint64_t n, unsigned long long* status,
                        cudaStream_t stream) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    k_count_dirty<<<unsigned(blocks), 256, 0, stream>>>(dirty, n, status);
}

}  // namespace lphb
