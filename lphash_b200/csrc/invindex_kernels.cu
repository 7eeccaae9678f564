// build-p Part 3 on the device: see invindex_kernels.cuh.  Everything here is a map, a prefix sum or a
// scatter over the distinct minimizers; the reference walks them one by one through four
// external_memory_vectors (src/partitioned_mphf.cpp:163-268).
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "device_mphf.cuh"
#include "invindex_kernels.cuh"

namespace lphb {

namespace {

unsigned grid_for(uint64_t n, unsigned per_block = 256) {
    uint64_t blocks = (n + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    return unsigned(blocks);
}

constexpr uint32_t kLeft = 0, kRightOrCollision = 1, kMaximal = 2, kNone = 3;  // include/quartet_wtree.hpp:7

// ref src/partitioned_mphf.cpp:96-100 (re-key) + :183-215 (type of a triplet)
__global__ void k_rekey(const __grid_constant__ DevPhf phf, const uint8_t* triplets, uint64_t n, uint32_t k, uint32_t m,
                        uint32_t* cells, unsigned long long* bad, unsigned long long* colliding) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint16_t* t = reinterpret_cast<const uint16_t*>(triplets + 10 * i);  // packed {u64 itself, u8 p1, u8 size}
        const uint64_t itself = uint64_t(t[0]) | (uint64_t(t[1]) << 16) | (uint64_t(t[2]) << 32) | (uint64_t(t[3]) << 48);
        const uint32_t p1 = t[4] & 0xFFu, size = t[4] >> 8;
        uint32_t type, a = 0, b = 0;
        if (size == 0) {
            type = kRightOrCollision;
            atomicAdd(colliding, 1ull);
        } else if (p1 == k - m) {
            if (size == k - m + 1) type = kMaximal;
            else type = kRightOrCollision, a = size;
        } else if (p1 == size - 1) {
            type = kLeft, a = p1 + 1;
        } else {
            type = kNone, a = size, b = p1;
        }
        const uint64_t order = phf_position(phf, murmur64(itself, phf.seed));
        if (order >= n) {
            atomicAdd(bad, 1ull);
            continue;
        }
        cells[order] = 0x80000000u | type | (a << 2) | (b << 10);
    }
}

// build-u (mphf_alt, ref src/unpartitioned_mphf.cpp:78-96): the triplet itself at its order, no types
__global__ void k_rekey_alt(const __grid_constant__ DevPhf phf, const uint8_t* triplets, uint64_t n, uint32_t* cells,
                            unsigned long long* bad) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint16_t* t = reinterpret_cast<const uint16_t*>(triplets + 10 * i);
        const uint64_t itself = uint64_t(t[0]) | (uint64_t(t[1]) << 16) | (uint64_t(t[2]) << 32) | (uint64_t(t[3]) << 48);
        const uint64_t order = phf_position(phf, murmur64(itself, phf.seed));
        if (order >= n) {
            atomicAdd(bad, 1ull);
            continue;
        }
        cells[order] = 0x80000000u | t[4];  // p1 | size << 8
    }
}
__global__ void k_split_alt(const uint32_t* cells, uint64_t n, uint8_t* p1, uint8_t* size, unsigned long long* unset) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t c = cells[i];
        if (!(c >> 31)) atomicAdd(unset, 1ull);
        p1[i] = uint8_t(c);
        size[i] = uint8_t(c >> 8);
    }
}

// one-hot of a cell's class in four 16-bit lanes: left, rc, none, msb
__device__ __forceinline__ uint64_t class_lanes(uint32_t cell) {
    if (!(cell >> 31)) return 0;
    const uint32_t type = cell & 3u;
    return (type == kLeft ? 1ull : 0) | (type == kRightOrCollision ? 1ull << 16 : 0) | (type == kNone ? 1ull << 32 : 0) |
           ((type >> 1) ? 1ull << 48 : 0);
}

__global__ void __launch_bounds__(256) k_cell_counts(const uint32_t* cells, uint64_t n, InvCounts* blk,
                                                     unsigned long long* unset) {
    using Reduce = cub::BlockReduce<uint64_t, 256>;
    __shared__ typename Reduce::TempStorage tmp;
    __shared__ typename Reduce::TempStorage tmp2;
    const uint64_t base = uint64_t(blockIdx.x) * kInvBlock;
    uint64_t lanes = 0, missing = 0;
    for (uint32_t j = 0; j < kInvBlock / 256; ++j) {
        const uint64_t i = base + j * 256 + threadIdx.x;
        if (i < n) {
            const uint32_t c = cells[i];
            lanes += class_lanes(c);
            missing += (c >> 31) ? 0 : 1;
        }
    }
    const uint64_t sum = Reduce(tmp).Sum(lanes);
    const uint64_t miss = Reduce(tmp2).Sum(missing);
    if (threadIdx.x == 0) {
        blk[blockIdx.x] = InvCounts{uint32_t(sum & 0xFFFF), uint32_t((sum >> 16) & 0xFFFF), uint32_t((sum >> 32) & 0xFFFF),
                                    uint32_t(sum >> 48)};
        if (miss) atomicAdd(unset, (unsigned long long)miss);
    }
}

struct CountsSum {
    __host__ __device__ InvCounts operator()(InvCounts const& x, InvCounts const& y) const {
        return InvCounts{x.left + y.left, x.rc + y.rc, x.none + y.none, x.msb + y.msb};
    }
};

// ref src/partitioned_mphf.cpp:183-215 (the four lists) + src/quartet_wtree.cpp:13-41 (the three bit vectors)
__global__ void __launch_bounds__(256) k_place(const uint32_t* cells, uint64_t n, const InvCounts* incl, uint64_t rs,
                                               uint64_t ns, uint64_t np, uint32_t* root, uint32_t* left_right,
                                               uint32_t* max_none, uint8_t* vals) {
    using Scan = cub::BlockScan<uint64_t, 256>;
    __shared__ typename Scan::TempStorage tmp;
    const uint64_t base = uint64_t(blockIdx.x) * kInvBlock;
    InvCounts before = blockIdx.x ? incl[blockIdx.x - 1] : InvCounts{0, 0, 0, 0};
    uint64_t running = 0;  // lanes of the cells of this block already passed
    for (uint32_t j = 0; j < kInvBlock / 256; ++j) {
        const uint64_t i = base + j * 256 + threadIdx.x;
        const uint32_t c = i < n ? cells[i] : 0;
        const uint64_t mine = class_lanes(c);
        uint64_t excl, total;
        Scan(tmp).ExclusiveSum(mine, excl, total);
        __syncthreads();  // tmp is reused by the next round
        excl += running;
        running += total;
        const uint32_t type = c & 3u, a = (c >> 2) & 0xFFu, b = (c >> 10) & 0xFFu;
        const bool live = c >> 31;
        const uint64_t r_left = before.left + (excl & 0xFFFF), r_rc = before.rc + ((excl >> 16) & 0xFFFF),
                       r_none = before.none + ((excl >> 32) & 0xFFFF), r_msb = before.msb + (excl >> 48);
        const unsigned rootbits = __ballot_sync(0xFFFFFFFFu, live && (type >> 1));
        if ((threadIdx.x & 31) == 0 && i < n) root[i >> 5] = rootbits;
        if (!live) continue;
        if (type == kLeft) {
            vals[r_left] = uint8_t(a);
        } else if (type == kRightOrCollision) {
            vals[rs + r_rc] = uint8_t(a);
            const uint64_t pos = i - r_msb;  // index among the cells with a clear root bit
            atomicOr(left_right + (pos >> 5), 1u << (pos & 31));
        } else if (type == kNone) {
            vals[ns + r_none] = uint8_t(a);
            vals[np + r_none] = uint8_t(b);
            atomicOr(max_none + (r_msb >> 5), 1u << (r_msb & 31));
        }
    }
}

struct U8ToU64 {
    __host__ __device__ uint64_t operator()(uint8_t v) const { return v; }
};

// position of the i-th one of the high bits: 0 for the leading zero value, then (v >> l) + i
// (ef_sequence.hpp:53, :68)
__device__ __forceinline__ uint64_t ef_one(const uint64_t* cum, uint64_t i, uint32_t l) {
    return i ? (cum[i - 1] >> l) + i : 0;
}

__global__ void k_ef_high(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t* high) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n_enc; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t pos = ef_one(cum, i, l);
        atomicOr(reinterpret_cast<unsigned long long*>(high) + (pos >> 6), 1ull << (pos & 63));
    }
}

// compact_vector of width l (pthash compact_vector.hpp:127-145): one thread per 64-bit word
__global__ void k_ef_low(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t* low, uint64_t low_words) {
    const uint64_t mask = (uint64_t(1) << l) - 1;
    for (uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; j < low_words; j += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t word = 0;
        if (l) {
            const uint64_t bit0 = j << 6;
            for (uint64_t i = bit0 / l; i < n_enc && i * l < bit0 + 64; ++i) {
                const uint64_t v = i ? (cum[i - 1] & mask) : 0;
                const uint64_t at = i * l;
                word |= at >= bit0 ? v << (at - bit0) : v >> (bit0 - at);
            }
        }
        low[j] = word;
    }
}

constexpr uint64_t kDBlock = 1024, kDSub = 32, kDMaxSpan = 1 << 16;  // darray.hpp:122-124

__global__ void k_darray_spans(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t n_blocks, uint64_t* sparse_cnt) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b <= n_blocks; b += uint64_t(gridDim.x) * blockDim.x) {
        if (b == n_blocks) {
            sparse_cnt[b] = 0;
            continue;
        }
        const uint64_t first = b * kDBlock, last = min(first + kDBlock, n_enc) - 1;
        const bool sparse = ef_one(cum, last, l) - ef_one(cum, first, l) >= kDMaxSpan;  // darray.hpp:101
        sparse_cnt[b] = sparse ? last - first + 1 : 0;
    }
}

__global__ void __launch_bounds__(256) k_darray_fill(const uint64_t* cum, uint64_t n_enc, uint32_t l, const uint64_t* sparse_cnt,
                                                     const uint64_t* ovf_off, int64_t* block_inventory,
                                                     uint16_t* subblock_inventory, uint64_t* overflow) {
    const uint64_t b = blockIdx.x, first = b * kDBlock, cnt = min(first + kDBlock, n_enc) - first;
    const bool sparse = sparse_cnt[b] != 0;
    const uint64_t front = ef_one(cum, first, l), off = ovf_off[b];
    if (threadIdx.x == 0) block_inventory[b] = sparse ? -int64_t(off) - 1 : int64_t(front);
    const uint64_t subs = (cnt + kDSub - 1) / kDSub;
    if (threadIdx.x < subs)
        subblock_inventory[b * (kDBlock / kDSub) + threadIdx.x] =
            sparse ? uint16_t(0xFFFF) : uint16_t(ef_one(cum, first + threadIdx.x * kDSub, l) - front);
    if (sparse)
        for (uint64_t t = threadIdx.x; t < cnt; t += blockDim.x) overflow[off + t] = ef_one(cum, first + t, l);
}

// rs_bit_vector::build_indices (include/rs_bit_vector.hpp:120-156), block = 8 words
__global__ void k_block_pop(const uint64_t* bits, uint64_t n_blocks, uint64_t* pop) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b <= n_blocks; b += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t c = 0;
        if (b < n_blocks)
            for (int j = 0; j < 8; ++j) c += __popcll(bits[b * 8 + j]);
        pop[b] = c;
    }
}
__global__ void k_rank_pairs(const uint64_t* bits, uint64_t n_blocks, const uint64_t* rank_before, uint64_t* pairs) {
    for (uint64_t b = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; b <= n_blocks; b += uint64_t(gridDim.x) * blockDim.x) {
        pairs[2 * b] = rank_before[b];
        uint64_t sub = 0;
        if (b < n_blocks) {
            uint64_t c = 0;
            for (int j = 0; j < 7; ++j) {
                c += __popcll(bits[b * 8 + j]);
                sub = (sub << 9) | c;
            }
        }
        pairs[2 * b + 1] = sub;
    }
}

}  // namespace

void launch_rekey(DevPhf const& phf, const uint8_t* triplets, uint64_t n, uint32_t k, uint32_t m, uint32_t* cells,
                  unsigned long long* bad, unsigned long long* colliding, cudaStream_t s) {
    if (!n) return;
    k_rekey<<<grid_for(n), 256, 0, s>>>(phf, triplets, n, k, m, cells, bad, colliding);
}

void launch_rekey_alt(DevPhf const& phf, const uint8_t* triplets, uint64_t n, uint32_t* cells, unsigned long long* bad,
                      cudaStream_t s) {
    if (n) k_rekey_alt<<<grid_for(n), 256, 0, s>>>(phf, triplets, n, cells, bad);
}
void launch_split_alt(const uint32_t* cells, uint64_t n, uint8_t* p1, uint8_t* size, unsigned long long* unset,
                      cudaStream_t s) {
    if (n) k_split_alt<<<grid_for(n), 256, 0, s>>>(cells, n, p1, size, unset);
}

void launch_cell_counts(const uint32_t* cells, uint64_t n, InvCounts* blk, unsigned long long* unset, cudaStream_t s) {
    if (!n) return;
    k_cell_counts<<<unsigned((n + kInvBlock - 1) / kInvBlock), 256, 0, s>>>(cells, n, blk, unset);
}

uint64_t inv_scan_tmp_bytes(uint64_t n_blocks) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, bytes, static_cast<InvCounts*>(nullptr), static_cast<InvCounts*>(nullptr),
                                   CountsSum{}, int(n_blocks));
    return bytes;
}
void launch_counts_scan(InvCounts* blk, uint64_t n_blocks, void* tmp, uint64_t tmp_bytes, cudaStream_t s) {
    if (!n_blocks) return;
    size_t bytes = tmp_bytes;
    cub::DeviceScan::InclusiveScan(tmp, bytes, blk, blk, CountsSum{}, int(n_blocks), s);
}

void launch_place(const uint32_t* cells, uint64_t n, const InvCounts* incl, uint64_t rs, uint64_t ns, uint64_t np,
                  uint32_t* root, uint32_t* left_right, uint32_t* max_none, uint8_t* vals, cudaStream_t s) {
    if (!n) return;
    k_place<<<unsigned((n + kInvBlock - 1) / kInvBlock), 256, 0, s>>>(cells, n, incl, rs, ns, np, root, left_right,
                                                                     max_none, vals);
}

uint64_t cum_tmp_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::TransformInputIterator<uint64_t, U8ToU64, const uint8_t*> in(nullptr, U8ToU64{});
    cub::DeviceScan::InclusiveSum(nullptr, bytes, in, static_cast<uint64_t*>(nullptr), int64_t(n));
    return bytes;
}
void launch_cumulative(const uint8_t* vals, uint64_t n, uint64_t* cum, void* tmp, uint64_t tmp_bytes, cudaStream_t s) {
    if (!n) return;
    size_t bytes = tmp_bytes;
    cub::TransformInputIterator<uint64_t, U8ToU64, const uint8_t*> in(vals, U8ToU64{});
    cub::DeviceScan::InclusiveSum(tmp, bytes, in, cum, int64_t(n), s);
}

void launch_ef_encode(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t* high, uint64_t* low, uint64_t low_words,
                      cudaStream_t s) {
    k_ef_high<<<grid_for(n_enc), 256, 0, s>>>(cum, n_enc, l, high);
    k_ef_low<<<grid_for(low_words), 256, 0, s>>>(cum, n_enc, l, low, low_words);
}

uint64_t darray_tmp_bytes(uint64_t n_blocks) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                  int64_t(n_blocks + 1));
    return bytes;
}
void launch_darray(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t n_blocks, uint64_t* sparse_cnt,
                   uint64_t* ovf_off, void* tmp, uint64_t tmp_bytes, cudaStream_t s) {
    k_darray_spans<<<grid_for(n_blocks + 1), 256, 0, s>>>(cum, n_enc, l, n_blocks, sparse_cnt);
    size_t bytes = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, bytes, sparse_cnt, ovf_off, int64_t(n_blocks + 1), s);
}
void launch_darray_fill(const uint64_t* cum, uint64_t n_enc, uint32_t l, uint64_t n_blocks, const uint64_t* sparse_cnt,
                        const uint64_t* ovf_off, int64_t* block_inventory, uint16_t* subblock_inventory,
                        uint64_t* overflow, cudaStream_t s) {
    if (!n_blocks) return;
    k_darray_fill<<<unsigned(n_blocks), 256, 0, s>>>(cum, n_enc, l, sparse_cnt, ovf_off, block_inventory,
                                                     subblock_inventory, overflow);
}

uint64_t rank_tmp_bytes(uint64_t n_blocks) { return darray_tmp_bytes(n_blocks); }
void launch_rank_pairs(const uint64_t* bits, uint64_t n_blocks, uint64_t* pop, uint64_t* pairs, void* tmp,
                       uint64_t tmp_bytes, cudaStream_t s) {
    k_block_pop<<<grid_for(n_blocks + 1), 256, 0, s>>>(bits, n_blocks, pop);
    size_t bytes = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, bytes, pop, pop, int64_t(n_blocks + 1), s);  // in place: rank before each block
    k_rank_pairs<<<grid_for(n_blocks + 1), 256, 0, s>>>(bits, n_blocks, pop, pairs);
}

}  // namespace lphb
