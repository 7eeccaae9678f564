// Compact form of the code stream for the trip over PCIe (lphb_query_stream_runs).
//
// Consecutive k-mers of one super-k-mer get consecutive codes (include/partitioned_mphf.hpp:131-145:
// local_rank is incremented or decremented by one), so the stream of 8-byte codes the reference returns
// is, by construction, a sequence of runs first, first+-1, first+-2, ...  This file re-expresses a code
// stream that is already on the device as 12-byte run records {u64 first, i32 n} (n > 0 ascending,
// n < 0 descending, |n| codes) - about 2 bytes per k-mer instead of 8 - and the host side
// (lphb_expand_runs, c_api.cu) turns them back into the identical uint64_t vector.  It works on any
// u64 stream (members, non-members, fallback codes, the non-ACGT quirk's spurious entries): where the
// stream is not an arithmetic progression the runs are simply short.
//
// Break rule (local, so every position decides on its own): with L_i = code[i] - code[i-1], a run
// starts at i iff i starts a contig, or L_i is not +-1, or the link before it is an unbroken +-1 of
// the other sign.  Inside a run all links are therefore equal.
#include <cub/device/device_scan.cuh>

#include "query_kernels.cuh"

namespace lphb {

namespace {

__global__ void k_run_starts(const uint64_t* code_off, uint64_t n_contigs, uint64_t o0, uint64_t n, uint8_t* start) {
    uint64_t c = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (c >= n_contigs) return;
    uint64_t a = code_off[c], b = code_off[c + 1];
    if (b > a && a >= o0 && a - o0 < n) start[a - o0] = 1;
}

// unbroken link between i-1 and i?  (i >= 1, neither... see file comment)
__device__ __forceinline__ int link_sign(const uint64_t* codes, uint64_t i) {
    const uint64_t d = codes[i] - codes[i - 1];
    return d == 1 ? 1 : (d == ~uint64_t(0) ? -1 : 0);
}

__global__ void __launch_bounds__(256) k_run_heads(const uint64_t* codes, uint64_t n, const uint8_t* start, uint8_t* head) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        bool h = i == 0 || start[i];
        if (!h) {
            const int s = link_sign(codes, i);
            h = s == 0;
            if (!h && i >= 2 && !start[i - 1]) {
                const int sp = link_sign(codes, i - 1);
                h = sp != 0 && sp != s;
            }
        }
        head[i] = h ? 1 : 0;
    }
}

struct U8ToU32 {
    const uint8_t* p;
    __host__ __device__ uint32_t operator()(uint64_t i) const { return p[i]; }
};

__global__ void __launch_bounds__(256) k_run_positions(const uint8_t* head, const uint32_t* rank, uint64_t n, uint32_t* head_at,
                                                       unsigned long long* n_runs_out) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        if (head[i]) head_at[rank[i]] = uint32_t(i);
        if (i == n - 1) {
            const uint32_t total = rank[i] + head[i];
            head_at[total] = uint32_t(n);
            *n_runs_out = total;
        }
    }
}

__global__ void __launch_bounds__(256) k_run_emit(const uint64_t* codes, const uint32_t* head_at, const unsigned long long* n_runs,
                                                  uint8_t* runs) {
    const uint64_t nr = *n_runs;
    for (uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; r < nr; r += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t a = head_at[r], b = head_at[r + 1];
        const uint64_t first = codes[a];
        int32_t len = int32_t(b - a);
        if (len > 1 && codes[a + 1] - first != 1) len = -len;
        // 12-byte records are only 4-byte aligned: three 32-bit stores
        uint32_t* dst = reinterpret_cast<uint32_t*>(runs + r * 12);
        dst[0] = uint32_t(first);
        dst[1] = uint32_t(first >> 32);
        dst[2] = uint32_t(len);
    }
}

}  // namespace

uint64_t runs_tmp_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint8_t*)nullptr, (uint32_t*)nullptr, n);
    return bytes + 256;
}

// codes[0..n) (n < 2^31) -> runs (12 B each, at most n), *d_n_runs = their number.  d_start/d_head: n + 8
// bytes each, d_rank: n + 1 u32, d_head_at: n + 1 u32 of scratch; code_off/n_contigs: where contigs start,
// in the same coordinates as `codes - o0`.
void launch_runs(const uint64_t* codes, uint64_t n, const uint64_t* code_off, uint64_t n_contigs, uint64_t o0,
                 uint8_t* d_start, uint8_t* d_head, uint32_t* d_rank, uint32_t* d_head_at, void* d_tmp,
                 uint64_t tmp_bytes, uint8_t* d_runs, unsigned long long* d_n_runs, cudaStream_t stream) {
    if (n == 0) {
        cudaMemsetAsync(d_n_runs, 0, sizeof(unsigned long long), stream);
        return;
    }
    cudaMemsetAsync(d_start, 0, n, stream);
    if (n_contigs) k_run_starts<<<unsigned((n_contigs + 255) / 256), 256, 0, stream>>>(code_off, n_contigs, o0, n, d_start);
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    k_run_heads<<<unsigned(blocks), 256, 0, stream>>>(codes, n, d_start, d_head);
    size_t bytes = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(d_tmp, bytes, d_head, d_rank, n, stream);
    k_run_positions<<<unsigned(blocks), 256, 0, stream>>>(d_head, d_rank, n, d_head_at, d_n_runs);
    k_run_emit<<<unsigned(blocks), 256, 0, stream>>>(codes, d_head_at, d_n_runs, d_runs);
}

}  // namespace lphb
