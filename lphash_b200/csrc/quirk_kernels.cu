// Contigs with non-ACGT bytes, run-parallel.
//
// The reference's streaming loop (include/partitioned_mphf.hpp:103-184) resets only its base counter
// and ring cursor at an invalid byte; the ring, the minimum's slot, and the rolling registers stay.
// What follows from the code:
//   (i)  inside a maximal ACGT run, every k-mer window gets exactly its stateless code (the first full
//        window rescans w fresh slots) - so all VALID k-mers of a dirty contig are ordinary work for
//        the tiled kernel, one "virtual contig" per run;
//   (ii) the only state-dependent outputs are the SPURIOUS codes pushed while m <= run < k, when the
//        cursor meets the stale minimum slot (:117-118, :146-171).  They depend on the state the
//        previous runs left, and that state is fully rewritten by any run of >= k bases (w fresh slots,
//        slot = m-mer index mod w, minimum = leftmost minimum of the last window).
// Hence: a dirty contig is cut into runs; the runs go through the tiled kernel as their own contigs;
// and one thread per EPISODE - a run of >= k bases, the shorter runs after it, and the first k-1 bases
// of the next long run - replays just those bases through the line-exact state machine (QuirkState
// below, the same code the one-thread-per-contig kernel of round 1 ran over whole contigs) to get the
// spurious codes.  The final sequence of a dirty contig is, run by run: its spurious codes, then its
// valid codes.  A record with one N in 5 Mbases costs two dozen sequential bases, not 5 million.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "device_mphf.cuh"
#include "query_kernels.cuh"

namespace lphb {

namespace {

// ---- line-exact restatement of the streaming loop's state ------------------------------------------
// ref: include/partitioned_mphf.hpp:73-184.  feed() is one iteration of its `for` loop for a valid
// base, invalid() its else-branch.  `push` is called with every code the reference pushes.
struct QuirkState {
    uint32_t k, m, w;
    uint64_t mm_mask, lo_mask, hi_mask;
    uint64_t ring_mm[64], ring_h[64];
    uint32_t cursor, min_slot, p1;
    uint64_t mmer, lo, hi, run;
    uint64_t g_rank, l_rank;  // last probe context (mm_context_t, partitioned_mphf.hpp:55-60)
    int32_t ctx_slope;        // slope of the last probe (0: colliding minimizer)

    __device__ void init(DevImage const& f) {
        k = f.k;
        m = f.m;
        w = f.w;
        mm_mask = (uint64_t(1) << (2 * m)) - 1;
        lo_mask = k >= 32 ? ~uint64_t(0) : ((uint64_t(1) << (2 * k)) - 1);
        hi_mask = k > 32 ? ((uint64_t(1) << (2 * k - 64)) - 1) : 0;
        for (uint32_t q = 0; q < w; ++q) ring_mm[q] = 0, ring_h[q] = 0;
        cursor = 0;
        min_slot = w;
        p1 = 0;
        mmer = lo = hi = run = 0;
        g_rank = l_rank = 0;
        ctx_slope = 1;
    }
    __device__ void invalid() {  // :179-183
        run = 0;
        cursor = 0;
    }
    // kProbe = false: only the state is advanced (codes are not wanted: no image access at all)
    template <bool kProbe, class Push>
    __device__ void feed(DevImage const& f, uint32_t code, Push&& push) {
        mmer = ((mmer << 2) | code) & mm_mask;
        hi = ((hi << 2) | (lo >> 62)) & hi_mask;
        lo = ((lo << 2) | code) & lo_mask;
        ++run;
        if (run < m) return;
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
                if (kProbe) {
                    if (ctx_slope == 0) l_rank = fallback_order(f, lo, hi);
                    else if (ctx_slope < 0) ++l_rank;  // RIGHT, NONE
                    else --l_rank;                     // LEFT, MAXIMAL
                    push(g_rank + l_rank);
                }
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
            if (kProbe) {
                Probe pr = probe_minimizer(f, ring_mm[min_slot]);
                ctx_slope = pr.slope;
                // split hval back into the reference's (global_rank, local_rank) so that the
                // +-1 continuation above reproduces :131-145 (mod 2^64)
                if (pr.slope == 0) { g_rank = pr.base; l_rank = fallback_order(f, lo, hi); }
                else if (pr.slope > 0) { g_rank = pr.base; l_rank = p1; }
                else { g_rank = pr.base; l_rank = uint64_t(0) - uint64_t(p1); }
                push(g_rank + l_rank);
            }
        }
        cursor = (cursor + 1) % w;
    }
};

// ---- segmented copy ----------------------------------------------------------------------------------------
// dst[i] = src(segment of i, i - seg_off[segment]) for i < seg_off[n_seg], with seg_off ascending (empty
// segments allowed).  A block takes 2048 consecutive elements, finds the segments of its first and last
// element once, and every thread searches only in between: a 5-Mbase contig is copied by 2400 blocks, a
// batch of 120-code reads costs four search steps per element.
constexpr int kSegChunk = 2048;
__device__ __forceinline__ uint64_t seg_of(const uint64_t* seg_off, uint64_t lo, uint64_t hi, uint64_t i) {
    while (hi - lo > 1) {  // seg_off[lo] <= i < seg_off[hi]
        const uint64_t mid = (lo + hi) >> 1;
        if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}
template <class T, class Src>
__global__ void __launch_bounds__(256) k_segcopy(T* dst, const uint64_t* seg_off, uint64_t n_seg, const unsigned long long* n_seg_dev,
                                                 Src src) {
    __shared__ uint64_t s_lo, s_hi;
    if (n_seg_dev) n_seg = *n_seg_dev;
    if (!n_seg) return;
    const uint64_t total = seg_off[n_seg];
    for (uint64_t c0 = uint64_t(blockIdx.x) * kSegChunk; c0 < total; c0 += uint64_t(gridDim.x) * kSegChunk) {
        const uint64_t c1 = c0 + kSegChunk < total ? c0 + kSegChunk : total;
        if (threadIdx.x == 0) {
            s_lo = seg_of(seg_off, 0, n_seg, c0);
            s_hi = seg_of(seg_off, 0, n_seg, c1 - 1) + 1;
        }
        __syncthreads();
        const uint64_t lo = s_lo, hi = s_hi;
        for (uint64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
            const uint64_t sg = seg_of(seg_off, lo, hi, i);
            dst[i] = src(sg, i - seg_off[sg]);
        }
        __syncthreads();
    }
}
template <class T, class Src>
void segcopy(T* dst, const uint64_t* seg_off, uint64_t n_seg, const unsigned long long* n_seg_dev, uint64_t total_hint, Src src,
             cudaStream_t stream) {
    uint64_t blocks = (total_hint + kSegChunk - 1) / kSegChunk;
    if (blocks < 1) blocks = 1;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    k_segcopy<T, Src><<<unsigned(blocks), 256, 0, stream>>>(dst, seg_off, n_seg, n_seg_dev, src);
}

struct ContigBytes {   // byte i of dirty contig j
    const char* bases;
    const uint64_t* offsets;
    const uint64_t* list;
    __device__ char operator()(uint64_t j, uint64_t i) const { return bases[offsets[list[j]] + i]; }
};
struct RunCodes {      // code i of run v: its spurious codes, then its valid ones
    const uint64_t* vcode_off;
    const uint64_t* vcodes;
    const uint64_t* spur;
    const uint32_t* spur_cnt;
    uint32_t w_cap;
    __device__ uint64_t operator()(uint64_t v, uint64_t i) const {
        const uint32_t ns = spur_cnt[v];
        return i < ns ? spur[v * w_cap + i] : vcodes[vcode_off[v] + (i - ns)];
    }
};
struct ContigCodes {   // code i of contig c: from the clean layout or from the non-ACGT scratch
    const uint64_t* src_a;
    const uint64_t* src_b;
    const uint64_t* src_off;
    const uint8_t* from_b;
    __device__ uint64_t operator()(uint64_t c, uint64_t i) const { return (from_b[c] ? src_b : src_a)[src_off[c] + i]; }
};

// flag[p] = 1 where a run (of valid or of invalid bytes) or a contig begins
__global__ void k_run_boundaries(const char* bases, uint64_t n, uint8_t* flag) {
    for (uint64_t p = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; p < n; p += uint64_t(gridDim.x) * blockDim.x) {
        if (p == 0) { flag[p] = 1; continue; }
        const bool a = nt4(uint8_t(bases[p - 1])) < 4, b = nt4(uint8_t(bases[p])) < 4;
        if (a != b) flag[p] = 1;  // (contig starts were set before: the array is zeroed, then marked)
    }
}
__global__ void k_mark_starts(const uint64_t* starts, uint64_t n_list, uint64_t n, uint8_t* flag) {
    uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (j < n_list && starts[j] < n) flag[starts[j]] = 1;
}
__global__ void k_close_offsets(uint64_t* voff, const unsigned long long* n_v, uint64_t n) { voff[*n_v] = n; }

// first run (virtual contig) of every dirty contig: its start is a boundary, so the search is exact
__global__ void k_first_run(const uint64_t* voff, const unsigned long long* n_v_p, const uint64_t* starts, uint64_t n_list,
                            uint64_t n, uint32_t* vstart) {
    const uint64_t n_v = *n_v_p;
    uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (j > n_list) return;
    const uint64_t target = j < n_list ? starts[j] : n;
    uint64_t lo = 0, hi = n_v;  // first v with voff[v] >= target
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (voff[mid] < target) lo = mid + 1; else hi = mid;
    }
    vstart[j] = uint32_t(lo);
}

// ---- episodes: the spurious codes ------------------------------------------------------------------------
// One thread per run v of >= k valid bases.  It rebuilds the state the run leaves (replaying its tail from a
// start that keeps the ring phase: a multiple of w m-mers before the end), then walks the following runs of
// its contig up to and including the first k-1 bases of the next run of >= k bases, recording what the
// reference pushes there.  spur[v * w_cap + i] / spur_cnt[v] per run.
__global__ void __launch_bounds__(64) k_episodes(const __grid_constant__ DevImage f, const char* bases, const uint64_t* voff,
                                                 const unsigned long long* n_v_p, const uint32_t* vstart, uint64_t n_list,
                                                 uint32_t w_cap, uint64_t* spur, uint32_t* spur_cnt) {
    const uint64_t n_v = *n_v_p;
    const uint64_t v = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (v >= n_v) return;
    const uint32_t k = f.k, m = f.m, w = f.w;
    const uint64_t s0 = voff[v], L = voff[v + 1] - s0;
    if (L < k || nt4(uint8_t(bases[s0])) > 3) return;  // not the head of an episode
    // last run of this run's contig
    uint64_t lo = 0, hi = n_list;  // vstart[lo] <= v < vstart[hi]
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (vstart[mid] <= v) lo = mid; else hi = mid;
    }
    const uint64_t v_end = vstart[lo + 1];
    if (v + 1 >= v_end) return;  // nothing after this run in its contig
    QuirkState st;
    st.init(f);
    auto drop = [](uint64_t) {};
    {
        // tail of the head run: start at a base index that is a multiple of w (slot = m-mer index mod w keeps
        // its phase) and leaves at least k + w bases, so that every slot and the minimum are the true ones
        uint64_t t0 = L > uint64_t(k + w) ? (L - (k + w)) / w * w : 0;
        for (uint64_t i = t0; i < L; ++i) st.feed<false>(f, nt4(uint8_t(bases[s0 + i])), drop);
    }
    for (uint64_t u = v + 1; u < v_end; ++u) {
        const uint64_t a = voff[u], n = voff[u + 1] - a;
        if (nt4(uint8_t(bases[a])) > 3) {  // a run of invalid bytes: each one resets counter and cursor
            st.invalid();
            continue;
        }
        const bool last = n >= k;                      // the next long run ends the episode ...
        const uint64_t upto = last ? uint64_t(k - 1) : n;  // ... after its first k-1 bases
        uint32_t cnt = 0;
        uint64_t* out = spur + u * w_cap;
        auto keep = [&](uint64_t code) { if (cnt < w_cap) out[cnt] = code; ++cnt; };
        for (uint64_t i = 0; i < upto; ++i) st.feed<true>(f, nt4(uint8_t(bases[a + i])), keep);
        spur_cnt[u] = cnt;
        if (last) break;
    }
}

// codes a run contributes: its spurious ones, then one per window of k valid bases
struct RunCount {
    const char* bases;
    const uint64_t* voff;
    const uint32_t* spur_cnt;
    uint32_t k;
    __host__ __device__ uint64_t operator()(uint64_t v) const {
        const uint64_t L = voff[v + 1] - voff[v];
        uint64_t n = spur_cnt[v];
#ifdef __CUDA_ARCH__
        if (L >= k && nt4(uint8_t(bases[voff[v]])) < 4) n += L - k + 1;
#endif
        return n;
    }
};
__global__ void k_close_counts(const char* bases, const uint64_t* voff, const uint32_t* spur_cnt, uint32_t k,
                               const unsigned long long* n_v_p, uint64_t* out_off) {
    const uint64_t n_v = *n_v_p;
    uint64_t total = 0;
    if (n_v) total = out_off[n_v - 1] + RunCount{bases, voff, spur_cnt, k}(n_v - 1);
    out_off[n_v] = total;
}

// per dirty contig: where its final sequence starts in `out`, and how many codes it has
__global__ void k_contig_counts(const uint64_t* out_off, const uint32_t* vstart, uint64_t n_list, uint64_t* q_off, uint64_t* counts) {
    uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (j >= n_list) return;
    q_off[j] = out_off[vstart[j]];
    counts[j] = out_off[vstart[j + 1]] - out_off[vstart[j]];
}


// ---- build side: minimizer::from_string over contigs with non-ACGT bytes (include/minimizer.hpp:138-151) ----
// An invalid byte flushes the open super-k-mer and restarts the window, so every maximal valid run is scanned like
// a contig of its own, with two exceptions the reference's loop has: m-mer ordinals (`id`) advance over valid runs
// only, and a run of exactly k bases that is FOLLOWED by an invalid byte counts its k-mer but never emits it (the
// first window is only searched when base k+1 arrives or the contig ends, minimizer.hpp:59-74, 153-162).  The scan
// kernels take a partition of the stream into contigs; the partition handed to them cuts invalid runs and those
// swallowed runs into pieces shorter than k, which hold no k-mer.
__global__ void k_shift_starts(const uint64_t* offsets, uint64_t n_contigs, uint64_t first, uint64_t* starts) {
    const uint64_t c = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (c <= n_contigs) starts[c] = offsets[c] - first;
}
__global__ void k_build_run_counts(const char* bases, const uint64_t* voff, const unsigned long long* n_v_p,
                                   const uint32_t* vstart, uint64_t n_list, uint32_t k, uint32_t m, uint64_t* pieces,
                                   uint64_t* ids, uint64_t* kmers) {
    const uint64_t n_v = *n_v_p;
    const uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (r > n_v) return;
    if (r == n_v) {
        pieces[r] = ids[r] = kmers[r] = 0;
        return;
    }
    const uint64_t L = voff[r + 1] - voff[r];
    const bool valid = nt4(uint8_t(bases[voff[r]])) < 4;
    bool swallowed = false;
    if (valid && L == k) {  // followed by an invalid byte of the same contig?
        uint64_t lo = 0, hi = n_list;  // first j with vstart[j] >= r + 1 (vstart[n_list] = n_v)
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (vstart[mid] < r + 1) lo = mid + 1; else hi = mid;
        }
        swallowed = vstart[lo] != r + 1;
    }
    const uint64_t step = k - 1;
    pieces[r] = (valid && !swallowed) ? 1 : (L + step - 1) / step;
    ids[r] = valid && L >= m ? L - m + 1 : 0;
    kmers[r] = valid && L >= k ? L - k + 1 : 0;
}
__global__ void k_build_pieces(const uint64_t* voff, const unsigned long long* n_v_p, uint32_t k, const uint64_t* pieces,
                               const uint64_t* ids, uint64_t mm_count_in, uint64_t n, uint64_t* poff, uint64_t* pid) {
    const uint64_t n_v = *n_v_p;
    const uint64_t r = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (r > n_v) return;
    if (r == n_v) {
        poff[pieces[r]] = n;
        pid[pieces[r]] = mm_count_in + ids[r];
        return;
    }
    const uint64_t first = pieces[r], cnt = pieces[r + 1] - first, step = k - 1;
    for (uint64_t i = 0; i < cnt; ++i) {
        poff[first + i] = voff[r] + i * step;
        pid[first + i] = mm_count_in + ids[r] + i * step;  // only read for single-piece (scanned) runs
    }
}

unsigned blocks_for(uint64_t n, unsigned cap = 148 * 32) {
    uint64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return unsigned(b);
}

}  // namespace

void launch_gather_contigs(const char* bases, const uint64_t* offsets, const uint64_t* list, const uint64_t* dst_off,
                           uint64_t n_list, uint64_t total_bytes, char* dst, cudaStream_t stream) {
    if (!n_list) return;
    segcopy<char>(dst, dst_off, n_list, nullptr, total_bytes, ContigBytes{bases, offsets, list}, stream);
}

uint64_t quirk_tmp_bytes(uint64_t n) {
    size_t a = 0, b = 0;
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::DeviceSelect::Flagged(nullptr, a, cnt, (const uint8_t*)nullptr, (uint64_t*)nullptr, (unsigned long long*)nullptr, n);
    RunCount op{nullptr, nullptr, nullptr, 1};
    cub::TransformInputIterator<uint64_t, RunCount, cub::CountingInputIterator<uint64_t>> it(cnt, op);
    cub::DeviceScan::ExclusiveSum(nullptr, b, it, (uint64_t*)nullptr, n);
    return (a > b ? a : b) + 256;
}

// runs of the gathered dirty contigs: voff[0..*n_v] (boundaries, closed with n), vstart[0..n_list] (first run per contig)
void launch_find_runs(const char* d_bases, uint64_t n, const uint64_t* d_starts, uint64_t n_list, uint8_t* d_flag,
                      uint64_t* d_voff, unsigned long long* d_n_v, uint32_t* d_vstart, void* d_tmp, uint64_t tmp_bytes,
                      cudaStream_t stream) {
    cudaMemsetAsync(d_flag, 0, n, stream);
    k_mark_starts<<<unsigned((n_list + 255) / 256), 256, 0, stream>>>(d_starts, n_list, n, d_flag);
    k_run_boundaries<<<blocks_for(n), 256, 0, stream>>>(d_bases, n, d_flag);
    size_t bytes = tmp_bytes;
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::DeviceSelect::Flagged(d_tmp, bytes, cnt, d_flag, d_voff, d_n_v, n, stream);
    k_close_offsets<<<1, 1, 0, stream>>>(d_voff, d_n_v, n);
    k_first_run<<<unsigned((n_list + 256) / 256), 256, 0, stream>>>(d_voff, d_n_v, d_starts, n_list, n, d_vstart);
}

// spurious codes of every run + the final per-run layout; n_v_host = number of runs (known on the host by now)
void launch_quirk_finish(DevImage const& img, const char* d_bases, const uint64_t* d_voff, uint64_t n_v_host,
                         const unsigned long long* d_n_v, const uint32_t* d_vstart, uint64_t n_list,
                         const uint64_t* d_vcode_off, const uint64_t* d_vcodes, uint64_t* d_spur, uint32_t* d_spur_cnt,
                         uint32_t w_cap, uint64_t* d_out_off, uint64_t* d_out, uint64_t* d_q_off, uint64_t* d_counts,
                         void* d_tmp, uint64_t tmp_bytes, cudaStream_t stream) {
    if (!n_v_host) return;
    cudaMemsetAsync(d_spur_cnt, 0, (n_v_host + 1) * 4, stream);
    k_episodes<<<unsigned((n_v_host + 63) / 64), 64, 0, stream>>>(img, d_bases, d_voff, d_n_v, d_vstart, n_list, w_cap, d_spur,
                                                                  d_spur_cnt);
    RunCount op{d_bases, d_voff, d_spur_cnt, img.k};
    cub::CountingInputIterator<uint64_t> cnt(0);
    cub::TransformInputIterator<uint64_t, RunCount, cub::CountingInputIterator<uint64_t>> it(cnt, op);
    size_t bytes = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(d_tmp, bytes, it, d_out_off, n_v_host, stream);
    k_close_counts<<<1, 1, 0, stream>>>(d_bases, d_voff, d_spur_cnt, img.k, d_n_v, d_out_off);
    k_contig_counts<<<unsigned((n_list + 255) / 256), 256, 0, stream>>>(d_out_off, d_vstart, n_list, d_q_off, d_counts);
    (void)d_out;
    (void)d_vcode_off;
    (void)d_vcodes;
}

void launch_quirk_emit(const uint64_t* d_vcode_off, const uint64_t* d_vcodes, const uint64_t* d_spur,
                       const uint32_t* d_spur_cnt, uint32_t w_cap, uint64_t n_v_host, const unsigned long long* d_n_v,
                       const uint64_t* d_out_off, uint64_t total_hint, uint64_t* d_out, cudaStream_t stream) {
    if (!n_v_host) return;
    segcopy<uint64_t>(d_out, d_out_off, n_v_host, d_n_v, total_hint, RunCodes{d_vcode_off, d_vcodes, d_spur, d_spur_cnt, w_cap},
                      stream);
}

// dst[dst_off[c] .. dst_off[c+1]) = src_c[src_off[c] ..) where src_c = from_b[c] ? src_b : src_a
void launch_assemble(uint64_t* dst, const uint64_t* dst_off, const uint64_t* src_a, const uint64_t* src_b,
                     const uint64_t* src_off, const uint8_t* from_b, uint64_t n_contigs, uint64_t total, cudaStream_t stream) {
    if (!n_contigs || !total) return;
    segcopy<uint64_t>(dst, dst_off, n_contigs, nullptr, total, ContigCodes{src_a, src_b, src_off, from_b}, stream);
}

void launch_shift_starts(const uint64_t* d_offsets, uint64_t n_contigs, uint64_t first, uint64_t* d_starts, cudaStream_t stream) {
    k_shift_starts<<<unsigned((n_contigs + 256) / 256), 256, 0, stream>>>(d_offsets, n_contigs, first, d_starts);
}

// per run: pieces / m-mer ordinals / k-mers, then exclusive sums in place (entry n_v = totals)
void launch_build_run_counts(const char* d_bases, const uint64_t* d_voff, const unsigned long long* d_n_v,
                             const uint32_t* d_vstart, uint64_t n_list, uint64_t n_v_host, uint32_t k, uint32_t m,
                             uint64_t* d_pieces, uint64_t* d_ids, uint64_t* d_kmers, void* d_tmp, uint64_t tmp_bytes,
                             cudaStream_t stream) {
    k_build_run_counts<<<unsigned((n_v_host + 256) / 256), 256, 0, stream>>>(d_bases, d_voff, d_n_v, d_vstart, n_list, k, m,
                                                                             d_pieces, d_ids, d_kmers);
    for (uint64_t* a : {d_pieces, d_ids, d_kmers}) {
        size_t bytes = tmp_bytes;
        cub::DeviceScan::ExclusiveSum(d_tmp, bytes, a, a, n_v_host + 1, stream);
    }
}

void launch_build_pieces(const uint64_t* d_voff, const unsigned long long* d_n_v, uint64_t n_v_host, uint32_t k,
                         const uint64_t* d_pieces, const uint64_t* d_ids, uint64_t mm_count_in, uint64_t n, uint64_t* d_poff,
                         uint64_t* d_pid, cudaStream_t stream) {
    k_build_pieces<<<unsigned((n_v_host + 256) / 256), 256, 0, stream>>>(d_voff, d_n_v, k, d_pieces, d_ids, mm_count_in, n,
                                                                         d_poff, d_pid);
}

}  // namespace lphb
