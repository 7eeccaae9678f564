// Random-gather throughput on sm_100a (B200): how many independent 4/8-byte loads at random
// addresses of a table the chip completes per second, as a function of table size (L2-resident or
// not), warps per SM, loads in flight per thread, and load path (ld.global.nc through L1,
// ld.global.cg around it, cp.async into shared memory).  With and without a concurrent write stream
// (the query kernel writes 8 B per k-mer while it gathers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: ld.global.nc.u64   1: ld.global.cg.u64 (L1 bypass)   2: ld.global.nc.u32   3: cp.async 8 B -> smem
// 4: dependent pair (u64 gather -> index of a u32 gather in a second table), like a PTHash probe
template <int MODE, int U>
__global__ void k_gather(const uint64_t* __restrict__ tab, uint64_t mask, const uint32_t* __restrict__ tab2,
                         uint64_t mask2, int iters, uint64_t* sink, uint64_t* wr, uint64_t wr_stride) {
    extern __shared__ uint64_t smem[];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(tid * 2654435761u + 12345u);
    uint64_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = mix(s + 0x9e3779b9u);
            idx[u] = (uint64_t(s) * 0x9E3779B97F4A7C15ull >> 20) & mask;
        }
        if constexpr (MODE == 0 || MODE == 1 || MODE == 2) {
            uint64_t v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (MODE == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v[u]) : "l"(tab + idx[u]));
                if (MODE == 1) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v[u]) : "l"(tab + idx[u]));
                if (MODE == 2) {
                    uint32_t w;
                    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(w) : "l"(reinterpret_cast<const uint32_t*>(tab) + idx[u]));
                    v[u] = w;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        } else if constexpr (MODE == 3) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t dst = uint32_t(__cvta_generic_to_shared(&smem[(u * blockDim.x + threadIdx.x)]));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(tab + idx[u]));
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int u = 0; u < U; ++u) acc += smem[u * blockDim.x + threadIdx.x];
        } else if constexpr (MODE == 5) {
            // per-lane 16-byte TMA bulk copies completing on one mbarrier per warp (off the LSU data path)
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            uint64_t* bar = &smem[(U * 2) * blockDim.x + warp];
            const uint32_t bar_a = uint32_t(__cvta_generic_to_shared(bar));
            if (it == 0) {
                if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
                __syncwarp();
            }
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(uint32_t(U * 32 * 16)) : "memory");
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t dst = uint32_t(__cvta_generic_to_shared(&smem[(u * blockDim.x + threadIdx.x) * 2]));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             ::"r"(dst), "l"(tab + (idx[u] & ~1ull)), "r"(bar_a) : "memory");
            }
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar_a), "r"(uint32_t(it & 1)) : "memory");
#pragma unroll
            for (int u = 0; u < U; ++u) acc += smem[(u * blockDim.x + threadIdx.x) * 2];
            __syncwarp();
        } else {
            uint64_t v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v[u]) : "l"(tab + idx[u]));
            uint32_t w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint64_t j = ((v[u] ^ idx[u]) * 0x9E3779B97F4A7C15ull >> 20) & mask2;
                asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(w[u]) : "l"(tab2 + j));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += w[u];
        }
        if (wr) {  // concurrent streaming writes: 8 B per gather, coalesced
#pragma unroll
            for (int u = 0; u < U; ++u)
                __stcs(reinterpret_cast<unsigned long long*>(wr) + (uint64_t(it) * U + u) * wr_stride + tid, acc + u);
        }
    }
    if (acc == 0x123456789abcdefull) sink[0] = acc;
}

template <int MODE, int U>
double run(const uint64_t* tab, uint64_t n, const uint32_t* tab2, uint64_t n2, int ctas, int threads, int iters,
           uint64_t* sink, uint64_t* wr) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    size_t smem = MODE == 3 ? size_t(U) * threads * 8 : (MODE == 5 ? size_t(U) * threads * 16 + 256 : 0);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_gather<MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    uint64_t stride = uint64_t(ctas) * threads;
    k_gather<MODE, U><<<ctas, threads, smem>>>(tab, n - 1, tab2, n2 - 1, 2, sink, wr, stride);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_gather<MODE, U><<<ctas, threads, smem>>>(tab, n - 1, tab2, n2 - 1, iters, sink, wr, stride);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    double gathers = double(ctas) * threads * iters * U * (MODE == 4 ? 2 : 1);
    return gathers / (best * 1e-3) / 1e9;  // G gathers / s
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint64_t maxn = 1ull << 28;  // 2 GiB of u64
    uint64_t *tab, *sink, *wr;
    uint32_t* tab2;
    cudaMalloc(&tab, maxn * 8);
    cudaMalloc(&tab2, maxn * 4);
    cudaMalloc(&sink, 64);
    cudaMalloc(&wr, 1ull << 31);
    cudaMemset(tab, 1, maxn * 8);
    cudaMemset(tab2, 1, maxn * 4);
    printf("# SMs %d.  G gathers/s (independent random loads; pairs count 2)\n", sms);
    printf("%-28s %8s %6s %4s %10s\n", "mode", "table", "w/SM", "U", "Ggather/s");
    for (uint64_t n : {1ull << 21, 1ull << 23}) {
        for (int w : {8, 16, 32}) {
            int threads = 128, ctas = sms * w * 32 / threads, iters = 128;
            double mb = n * 8 / 1048576.0;
            printf("%-28s %6.0fMB %6d %4d %10.1f\n", "ld.nc.u64", mb, w, 3, run<0, 3>(tab, n, tab2, n, ctas, threads, iters, sink, nullptr));
            printf("%-28s %6.0fMB %6d %4d %10.1f\n", "tma bulk 16B", mb, w, 1, run<5, 1>(tab, n, tab2, n, ctas, threads, iters, sink, nullptr));
            printf("%-28s %6.0fMB %6d %4d %10.1f\n", "tma bulk 16B", mb, w, 3, run<5, 3>(tab, n, tab2, n, ctas, threads, iters, sink, nullptr));
            printf("%-28s %6.0fMB %6d %4d %10.1f\n", "tma bulk 16B", mb, w, 6, run<5, 6>(tab, n, tab2, n, ctas, threads, iters, sink, nullptr));
            printf("%-28s %6.0fMB %6d %4d %10.1f\n", "cp.async 8B -> smem", mb, w, 6, run<3, 6>(tab, n, tab2, n, ctas, threads, iters, sink, nullptr));
            fflush(stdout);
        }
    }
    return 0;
}
