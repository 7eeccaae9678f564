// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin extern "C" harness around the UNMODIFIED reference (jermp/lphash @ 9af2242, pthash @
// 561e513).  It is compiled by oracle/build_ref.sh together with the reference's own sources,
// taken where they lie under $LPHASH_REF_DIR (/root/reference), into
// oracle/_ref/libref{64,128}.so.  Nothing of the reference is copied into this repository; this
// file only *calls* the reference's public entry points:
//
//   lphash::mphf::build                     src/partitioned_mphf.cpp:33-145
//   lphash::mphf::operator()                include/partitioned_mphf.hpp:73-197
//   lphash::minimizer::from_string          include/minimizer.hpp:11-170
//   lphash::minimizer::classify             src/minimizer.cpp:5-50
//   lphash::minimizer::get_colliding_kmers  include/minimizer.hpp:172-319
//   lphash::mphf_alt::build / operator()    src/unpartitioned_mphf.cpp:31-140, include/unpartitioned_mphf.hpp:72-192
//   essentials::save / load                 pthash/external/essentials/include/essentials.hpp:595-607
//   pthash::single_phf::build_in_external_memory  (as mphf::build_minimizers_mphf calls it, src/partitioned_mphf.cpp:147-153)
//   lphash::ef_sequence::encode             include/ef_sequence.hpp:36-75
//
// Users: tests/ (to pin the CPU restatement in oracle/lphash_oracle.cpp and to generate the
// golden fixtures under tests/golden/), bench.py's cpu_baseline / --impl reference legs, and the
// fixture builder (the reference's build-p is the only producer of `.lph` files; PTHash
// construction is out of scope of the GPU path).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "include/partitioned_mphf.hpp"
#include "include/unpartitioned_mphf.hpp"
#include "include/minimizer.hpp"

using lphash::kmer_t;

namespace {
thread_local std::string g_err;
int fail(std::exception const& e) {
    g_err = e.what();
    return -1;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_kmer_bits() { return int(sizeof(kmer_t) * 8); }

// build-p: same calls as src/build.cpp:23-30 (build, then essentials::save).
int ref_build(const char* input, int k, int m, uint64_t seed, double c, int threads,
              int max_memory_gb, const char* tmp_dir, const char* output, int verbose,
              char* csv_out, uint64_t csv_cap) {
    try {
        lphash::configuration config;
        config.input_filename = input;
        config.output_filename = output ? output : "";
        config.k = k;
        config.m = m;
        config.mm_seed = seed;
        config.c = c;
        config.num_threads = threads;
        config.max_memory = max_memory_gb;
        config.tmp_dirname = tmp_dir ? tmp_dir : ".";
        config.check = false;
        config.verbose = verbose != 0;
        lphash::mphf f;
        std::ostringstream csv;
        f.build(config, csv);
        if (output && output[0]) essentials::save(f, output);
        if (csv_out && csv_cap) {
            std::string s = csv.str();
            std::strncpy(csv_out, s.c_str(), csv_cap - 1);
            csv_out[csv_cap - 1] = 0;
        }
        return 0;
    } catch (std::exception const& e) { return fail(e); }
}

// build-u / query-u: the unpartitioned variant (lphash::mphf_alt), same calling conventions
int ref_build_alt(const char* input, int k, int m, uint64_t seed, double c, int threads,
                  int max_memory_gb, const char* tmp_dir, const char* output, char* csv_out, uint64_t csv_cap) {
    try {
        lphash::configuration config;
        config.input_filename = input;
        config.output_filename = output ? output : "";
        config.k = k;
        config.m = m;
        config.mm_seed = seed;
        config.c = c;
        config.num_threads = threads;
        config.max_memory = max_memory_gb;
        config.tmp_dirname = tmp_dir ? tmp_dir : ".";
        config.check = false;
        config.verbose = false;
        lphash::mphf_alt f;
        std::ostringstream csv;
        f.build(config, csv);
        if (output && output[0]) essentials::save(f, output);
        if (csv_out && csv_cap) {
            std::string s = csv.str();
            std::strncpy(csv_out, s.c_str(), csv_cap - 1);
            csv_out[csv_cap - 1] = 0;
        }
        return 0;
    } catch (std::exception const& e) { return fail(e); }
}
void* ref_load_alt(const char* path) {
    try {
        auto* f = new lphash::mphf_alt();
        essentials::load(*f, path);
        return f;
    } catch (std::exception const& e) {
        fail(e);
        return nullptr;
    }
}
void ref_free_alt(void* h) { delete static_cast<lphash::mphf_alt*>(h); }
uint64_t ref_kmer_count_alt(void* h) { return static_cast<lphash::mphf_alt*>(h)->get_kmer_count(); }
int64_t ref_query_alt(void* h, const char* contig, uint64_t len, int streaming, uint64_t* out, uint64_t cap) {
    try {
        auto const& f = *static_cast<lphash::mphf_alt const*>(h);
        auto codes = f(contig, len, streaming != 0);
        uint64_t n = codes.size() < cap ? codes.size() : cap;
        if (out && n) std::memcpy(out, codes.data(), n * sizeof(uint64_t));
        return int64_t(codes.size());
    } catch (std::exception const& e) { return fail(e); }
}

void* ref_load(const char* path) {
    try {
        auto* f = new lphash::mphf();
        essentials::load(*f, path);
        return f;
    } catch (std::exception const& e) {
        fail(e);
        return nullptr;
    }
}

void ref_free(void* h) { delete static_cast<lphash::mphf*>(h); }

uint64_t ref_kmer_count(void* h) { return static_cast<lphash::mphf*>(h)->get_kmer_count(); }
uint64_t ref_minimizer_count(void* h) { return static_cast<lphash::mphf*>(h)->get_minimizer_L0(); }

// hf(contig, len, streaming): returns the number of codes the reference produced; copies at most
// `cap` of them.  (A negative return = exception.)
int64_t ref_query(void* h, const char* contig, uint64_t len, int streaming, uint64_t* out,
                  uint64_t cap) {
    try {
        auto const& f = *static_cast<lphash::mphf const*>(h);
        auto codes = f(contig, len, streaming != 0);
        uint64_t n = codes.size() < cap ? codes.size() : cap;
        if (out && n) std::memcpy(out, codes.data(), n * sizeof(uint64_t));
        return int64_t(codes.size());
    } catch (std::exception const& e) { return fail(e); }
}

// The in-memory multi-thread CPU baseline of BASELINE.md §3(ii): `threads` std::threads call the
// const operator() over disjoint contiguous contig ranges (balanced by bases).  If `out` is
// non-null the codes are written at out + out_offsets[c] (out_offsets must then hold the
// exclusive scan of the per-contig counts for clean input, i.e. max(0, L-k+1)).  Returns the
// wall-clock seconds of the parallel region; *total gets the number of codes produced.
double ref_query_batch(void* h, const char* bases, const uint64_t* offsets, uint64_t n_contigs,
                       int threads, uint64_t* out, const uint64_t* out_offsets, uint64_t* total) {
    auto const& f = *static_cast<lphash::mphf const*>(h);
    if (threads < 1) threads = 1;
    std::vector<uint64_t> cut(threads + 1, n_contigs);
    cut[0] = 0;
    uint64_t total_bases = offsets[n_contigs] - offsets[0];
    for (int t = 1; t < threads; ++t) {
        uint64_t target = offsets[0] + total_bases * t / threads;
        uint64_t lo = cut[t - 1], hi = n_contigs;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (offsets[mid] < target) lo = mid + 1; else hi = mid;
        }
        cut[t] = lo;
    }
    std::vector<uint64_t> counts(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            uint64_t n = 0;
            for (uint64_t c = cut[t]; c < cut[t + 1]; ++c) {
                auto codes = f(bases + offsets[c], offsets[c + 1] - offsets[c], true);
                n += codes.size();
                if (out && !codes.empty())
                    std::memcpy(out + out_offsets[c], codes.data(), codes.size() * sizeof(uint64_t));
                essentials::do_not_optimize_away(codes.data());
            }
            counts[t] = n;
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    uint64_t n = 0;
    for (auto c : counts) n += c;
    if (total) *total = n;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Build-side scan.  Opaque accumulator = the reference's external_memory_vector<mm_record_t>.
// The comparator orders by id (= scan order, ids are assigned monotonically, minimizer.hpp:59),
// so dumping it gives the stream in the order from_string emitted it.
struct scan_acc {
    lphash::external_memory_vector<lphash::mm_record_t> vec;
    scan_acc(std::string const& tmp, bool by_minimizer, uint64_t bytes = uint64_t(4) * essentials::GB)
        : vec(bytes,
              by_minimizer
                  ? std::function<bool(lphash::mm_record_t const&, lphash::mm_record_t const&)>(
                        [](auto const& a, auto const& b) { return a.itself < b.itself; })
                  : std::function<bool(lphash::mm_record_t const&, lphash::mm_record_t const&)>(
                        [](auto const& a, auto const& b) { return a.id < b.id; }),
              tmp, lphash::get_group_id()) {}
};

void* ref_scan_new(const char* tmp_dir, int by_minimizer) {
    try {
        return new scan_acc(tmp_dir ? tmp_dir : ".", by_minimizer != 0);
    } catch (std::exception const& e) {
        fail(e);
        return nullptr;
    }
}
void ref_scan_free(void* a) { delete static_cast<scan_acc*>(a); }

// from_string on one contig; mm_count is the in/out running m-mer ordinal.  Returns #k-mers.
uint64_t ref_from_string(void* a, const char* contig, uint64_t len, uint32_t k, uint32_t m,
                         uint64_t seed, uint64_t* mm_count) {
    auto* acc = static_cast<scan_acc*>(a);
    return lphash::minimizer::from_string<pthash::murmurhash2_64>(contig, len, k, m, seed, false,
                                                                  *mm_count, acc->vec);
}
uint64_t ref_scan_size(void* a) { return static_cast<scan_acc*>(a)->vec.size(); }

// Dump the records (18-byte packed mm_record_t, constants.hpp:26-33) in comparator order.
uint64_t ref_scan_dump(void* a, void* out, uint64_t cap_records) {
    auto* acc = static_cast<scan_acc*>(a);
    uint64_t n = 0;
    auto* dst = static_cast<uint8_t*>(out);
    for (auto it = acc->vec.cbegin(); it != acc->vec.cend() && n < cap_records; ++it, ++n)
        std::memcpy(dst + n * sizeof(lphash::mm_record_t), &*it, sizeof(lphash::mm_record_t));
    return n;
}

// classify (src/minimizer.cpp:5-50) on an accumulator created with by_minimizer=1.
// Outputs: unique triplets (10-byte packed mm_triplet_t) and the ascending colliding ids.
int ref_classify(void* a, const char* tmp_dir, void* triplets_out, uint64_t trip_cap,
                 uint64_t* n_triplets, uint64_t* ids_out, uint64_t ids_cap, uint64_t* n_ids) {
    try {
        auto* acc = static_cast<scan_acc*>(a);
        auto res = lphash::minimizer::classify(acc->vec, 4, tmp_dir ? tmp_dir : ".");
        uint64_t n = 0;
        auto* dst = static_cast<uint8_t*>(triplets_out);
        for (auto it = res.first.cbegin(); it != res.first.cend(); ++it, ++n)
            if (n < trip_cap)
                std::memcpy(dst + n * sizeof(lphash::mm_triplet_t), &*it,
                            sizeof(lphash::mm_triplet_t));
        *n_triplets = n;
        n = 0;
        for (auto it = res.second.cbegin(); it != res.second.cend(); ++it, ++n)
            if (n < ids_cap) ids_out[n] = *it;
        *n_ids = n;
        return 0;
    } catch (std::exception const& e) { return fail(e); }
}

// get_colliding_kmers over a batch of contigs with a sorted id list; the id cursor and mm_count
// persist across contigs exactly as in mphf::build Part 4 (src/partitioned_mphf.cpp:120-129).
// k-mers are written as sizeof(kmer_t) little-endian bytes each.  Returns #k-mers.
int64_t ref_colliding_kmers(const char* bases, const uint64_t* offsets, uint64_t n_contigs,
                            uint32_t k, uint32_t m, uint64_t seed, const uint64_t* ids,
                            uint64_t n_ids, const char* tmp_dir, void* kmers_out,
                            uint64_t cap_kmers) {
    try {
        std::string tmp = tmp_dir ? tmp_dir : ".";
        lphash::external_memory_vector<uint64_t> idv(
            uint64_t(1) * essentials::GB, [](uint64_t x, uint64_t y) { return x < y; }, tmp,
            lphash::get_group_id() + "i");
        for (uint64_t i = 0; i < n_ids; ++i) idv.push_back(ids[i]);
        lphash::external_memory_vector<kmer_t, false> acc(uint64_t(4) * essentials::GB, tmp,
                                                          lphash::get_group_id() + "k");
        auto start = idv.cbegin();
        auto stop = idv.cend();
        uint64_t id = 0;
        for (uint64_t c = 0; c < n_contigs; ++c)
            lphash::minimizer::get_colliding_kmers<pthash::murmurhash2_64>(
                bases + offsets[c], offsets[c + 1] - offsets[c], k, m, seed, false, start, stop, id,
                acc);
        uint64_t n = 0;
        auto* dst = static_cast<uint8_t*>(kmers_out);
        for (auto it = acc.cbegin(); it != acc.cend(); ++it, ++n)
            if (n < cap_kmers) std::memcpy(dst + n * sizeof(kmer_t), &*it, sizeof(kmer_t));
        return int64_t(n);
    } catch (std::exception const& e) { return fail(e); }
}

// Timed multi-thread build-side scan baseline (from_string with per-thread accumulators).
double ref_scan_batch(const char* bases, const uint64_t* offsets, uint64_t n_contigs, uint32_t k,
                      uint32_t m, uint64_t seed, int threads, const char* tmp_dir,
                      uint64_t* total_kmers, uint64_t* total_records) {
    if (threads < 1) threads = 1;
    std::string tmp = tmp_dir ? tmp_dir : ".";
    std::vector<uint64_t> kc(threads, 0), rc(threads, 0);
    std::vector<std::unique_ptr<scan_acc>> accs;
    for (int t = 0; t < threads; ++t) accs.emplace_back(new scan_acc(tmp, false, uint64_t(512) << 20));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            uint64_t lo = n_contigs * t / threads, hi = n_contigs * (t + 1) / threads;
            uint64_t id = 0, n = 0;
            for (uint64_t c = lo; c < hi; ++c)
                n += lphash::minimizer::from_string<pthash::murmurhash2_64>(
                    bases + offsets[c], offsets[c + 1] - offsets[c], k, m, seed, false, id,
                    accs[t]->vec);
            kc[t] = n;
            rc[t] = accs[t]->vec.size();
        });
    }
    for (auto& th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    uint64_t a = 0, b = 0;
    for (int t = 0; t < threads; ++t) a += kc[t], b += rc[t];
    if (total_kmers) *total_kmers = a;
    if (total_records) *total_records = b;
    return std::chrono::duration<double>(t1 - t0).count();
}

// minimizer_order for an arbitrary key set: the configuration of src/partitioned_mphf.cpp:45-52 and the call of
// :147-153, saved with essentials::save to `output`.
int ref_build_minimizer_phf(const uint64_t* keys, uint64_t n, double c, int threads, const char* tmp_dir,
                            const char* output) {
    try {
        pthash::build_configuration cfg;
        cfg.minimal_output = true;
        cfg.seed = lphash::constants::default_pthash_seed;
        cfg.c = c;
        cfg.alpha = 0.94;
        cfg.verbose_output = false;
        cfg.num_threads = threads;
        cfg.ram = uint64_t(4) * essentials::GB;
        cfg.tmp_dir = tmp_dir ? tmp_dir : ".";
        lphash::pthash_minimizers_mphf_t f;
        f.build_in_external_memory(keys, n, cfg);
        essentials::save(f, output);
        return 0;
    } catch (std::exception const& e) { return fail(e); }
}

// ef_sequence::encode(values, n, u) (values already cumulative, as cumulative_iterator hands them over), saved
// with essentials::save to `output`.
int ref_ef_sequence(const uint64_t* values, uint64_t n, uint64_t universe, const char* output) {
    try {
        lphash::ef_sequence ef;
        ef.encode(values, n, universe);
        essentials::save(ef, output);
        return 0;
    } catch (std::exception const& e) { return fail(e); }
}

}  // extern "C"
