// Instruction-throughput microbenchmarks for sm_100a (B200): which pipe an integer op runs on and
// its issue rate.  Output: warp-instructions per cycle per SM for each op mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__device__ __forceinline__ void step(uint32_t (&a)[CHAINS], uint32_t (&b)[CHAINS], uint32_t c, uint32_t d) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        if constexpr (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
        if constexpr (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(c));
        if constexpr (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
        if constexpr (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
        if constexpr (OP == 4) {
            uint64_t t;
            asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a[i]), "r"(c), "l"((uint64_t(b[i]) << 32) | a[i]));
            a[i] = uint32_t(t); b[i] = uint32_t(t >> 32);
        }
        if constexpr (OP == 5) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
        if constexpr (OP == 6) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(d));
        if constexpr (OP == 7) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        if constexpr (OP == 8) a[i] = __vimin3_u32(a[i], b[i], c + i);
        if constexpr (OP == 9) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
        if constexpr (OP == 10) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
        if constexpr (OP == 11) {  // lop3 + imad alternating (two pipes)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(d));
        }
        if constexpr (OP == 12) a[i] = __shfl_down_sync(0xFFFFFFFFu, a[i], 1);
        if constexpr (OP == 13) {  // 2 lop3 + 1 imad
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(c), "r"(d));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
        }
        if constexpr (OP == 14) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a[i]));
        if constexpr (OP == 15) asm volatile("{.reg .pred p; setp.lt.u64 p, %1, %2; selp.u32 %0, %3, %0, p;}" : "+r"(a[i]) : "l"((uint64_t(b[i]) << 32) | a[i]), "l"((uint64_t(c) << 32) | d), "r"(d));
        if constexpr (OP == 16) asm volatile("bfe.u32 %0, %0, 3, 9;" : "+r"(a[i]));
        if constexpr (OP == 17) asm volatile("mul.lo.u32 %0, %0, 16;" : "+r"(a[i]));  // shift by IMAD imm
        if constexpr (OP == 18) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.u32 %0, %0, %1, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if constexpr (OP == 19) { a[i] = __vimin_s32_relu(a[i], b[i]); }
        if constexpr (OP == 20) { a[i] = __viaddmin_u32(a[i], c, b[i]); }
        if constexpr (OP == 21) {  // 64-bit multiply by a constant
            uint64_t t = (uint64_t(b[i]) << 32) | a[i];
            t *= 0xc6a4a7935bd1e995ULL;
            a[i] = uint32_t(t); b[i] = uint32_t(t >> 32);
        }
        if constexpr (OP == 22) {  // mul.wide (no accumulate)
            uint64_t t;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(c));
            a[i] = uint32_t(t) ^ uint32_t(t >> 32);
        }
        if constexpr (OP == 23) {  // xorshift 64: x ^= x >> 47
            uint64_t t = (uint64_t(b[i]) << 32) | a[i];
            t ^= t >> 47;
            a[i] = uint32_t(t); b[i] = uint32_t(t >> 32) + c;
        }
        if constexpr (OP == 24) {  // full murmur64 of a 64-bit word
            uint64_t v = (uint64_t(b[i]) << 32) | a[i];
            const uint64_t M = 0xc6a4a7935bd1e995ULL;
            uint64_t h = (uint64_t(c) << 32 | d) ^ (8 * M);
            uint64_t x = v * M; x ^= x >> 47; x *= M; h = (h ^ x) * M; h ^= h >> 47; h *= M; h ^= h >> 47;
            a[i] = uint32_t(h); b[i] = uint32_t(h >> 32);
        }
        if constexpr (OP == 25) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a[i]) : "r"(c));  // add through IMAD
        if constexpr (OP == 26) {  // lds.32
            extern __shared__ uint32_t sm[];
            a[i] = sm[(a[i] & 1023)];
        }
        if constexpr (OP == 27) {  // lop3 + imad + lds
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(d));
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t c, uint32_t d, long long* cyc) {
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x * 17 + i; b[i] = threadIdx.x ^ (i * 77); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) step<OP>(a, b, c, d);
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

void run_quiet() {
    static uint32_t* out = nullptr; static long long* cyc = nullptr;
    if (!out) { cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 148 * 4 * 8); }
    k<0><<<148 * 4, 256, 4096>>>(out, 1u, 2u, cyc);
    cudaDeviceSynchronize();
}

template <int OP>
void run(const char* name, int per_step) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * 4, threads = 256;  // 32 warps per SM = 8 per SMSP
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, size_t(blocks) * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 1e9f;
    for (int rep = 0; rep < 12; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<blocks, threads, 4096>>>(out, 12345u, 678u, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float t; cudaEventElapsedTime(&t, e0, e1);
        if (rep >= 4 && t < ms) ms = t;
    }
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double winst = double(ITERS) * CHAINS * per_step * 8.0 * 4;  // warp-instr per SM (32 warps)
    printf("%-30s ms %.4f  warp-inst/ns/SM %.3f  = %.3f per cycle per SM at 1.92 GHz\n", name, ms,
           winst / (ms * 1e6), winst / (ms * 1e6) / 1.92);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int i = 0; i < 300; ++i) run_quiet();
    run<0>("lop3", 1);
    run<1>("shf.l.wrap", 1);
    run<14>("shl imm", 1);
    run<2>("add.u32", 1);
    run<3>("mad.lo.u32", 1);
    run<17>("mul.lo imm16", 1);
    run<4>("mad.wide.u32", 1);
    run<5>("mad.hi.u32", 1);
    run<6>("setp.u32+selp", 2);
    run<15>("setp.u64+selp", 2);
    run<18>("selp", 1);
    run<7>("min.u32", 1);
    run<8>("vimin3_u32", 1);
    run<19>("vimin_s32_relu", 1);
    run<20>("viaddmin_u32", 1);
    run<9>("prmt", 1);
    run<10>("popc", 1);
    run<16>("bfe.u32", 1);
    run<11>("lop3+imad 1:1", 2);
    run<13>("lop3+lop3+imad 2:1", 3);
    run<12>("shfl.down", 1);
    run<21>("mul.lo.u64 const (3-4 inst)", 1);
    run<22>("mul.wide.u32 + lop", 2);
    run<23>("xorshift64 >>47 (+add)", 1);
    run<24>("murmur64 (count as 1)", 1);
    run<25>("imad a*1+c", 1);
    run<26>("lds.32 dependent", 1);
    return 0;
}
