// Host ingest for the query / build drivers: FASTA / FASTQ text -> the batch layout the C ABI consumes
// (concatenated bases + uint64 offsets[n_records + 1]).
//
// Replaces, for batched calls, the record loop of the reference's drivers
//     while (kseq_read(seq) >= 0) hf(seq->seq.s, seq->seq.l, true);            src/query.cpp:51-55
//     while (kseq_read(seq) >= 0) minimizer::from_string(seq->seq.s, ...)      src/partitioned_mphf.cpp:70-77
// and follows kseq.h's record grammar exactly, so that record i here is the i-th `seq->seq.s`:
//   * a record starts at the first '>' or '@'; the rest of that line (name, comment) is skipped;
//   * sequence lines follow until a line whose FIRST character is '>', '@' or '+' (markers are only
//     recognised at line starts); line ends ("\n", "\r\n") are removed, nothing else is altered
//     (case, N, any other byte are preserved);
//   * '+' opens a FASTQ quality section: the rest of that line is skipped, then quality lines are
//     consumed until they hold at least as many characters as the sequence (multi-line FASTQ works);
//   * after a FASTQ record, bytes up to the next '>' or '@' are skipped.
// kseq reads through gzread one byte-buffer at a time (~9 ns/base); this works on the whole text in
// memory with memchr (a few GB/s), which is what keeps the GPU path from waiting on the parser.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef LPHASH_B200_WITH_ZLIB
#include <zlib.h>
#endif

namespace lphash_b200 {
namespace fastx {

struct Batch {
    std::vector<char> bases;         // all records' sequences, concatenated
    std::vector<uint64_t> offsets;   // n_records + 1 (offsets[0] = 0)
    uint64_t n_records() const { return offsets.empty() ? 0 : offsets.size() - 1; }
};

// Appends the records of `data[0..n)` to `out` and returns how many bytes were consumed.
//   final = true : the text ends here (end of file): every record kseq_read would return is appended, n is returned.
//   final = false: more text follows.  A record is only trusted once the NEXT record's header line has been seen
//                  (a FASTA sequence or a FASTQ quality string may continue in the next chunk), so the last record
//                  in the chunk is held back: the return value is the position of its header character and the
//                  caller passes data[ret..n) again, followed by the next chunk.
// *stopped is set when kseq_read would have returned an error (FASTQ quality string of the wrong length, with more
// text after it): the reference's `while (kseq_read(seq) >= 0)` loop ends there, and so must the caller.
inline size_t parse_some(const char* data, size_t n, Batch& out, bool final, bool* stopped = nullptr,
                         uint64_t* added_out = nullptr) {
    if (out.offsets.empty()) out.offsets.push_back(0);
    if (stopped) *stopped = false;
    const char* p = data;
    const char* const end = data + n;
    uint64_t added = 0;
    size_t last_seq_begin = 0;          // where the bases of the last accepted record start in out.bases
    const char* last_rec_start = nullptr;  // its header character
    auto line_end = [&](const char* q) -> const char* {
        const void* e = q < end ? std::memchr(q, '\n', size_t(end - q)) : nullptr;
        return e ? static_cast<const char*>(e) : end;
    };
    auto done = [&](size_t consumed) {
        if (added_out) *added_out = added;
        return consumed;
    };
    // skip to the first header character
    while (p < end && *p != '>' && *p != '@') ++p;
    while (p < end) {
        // p is at a header character: skip the header line
        const char* const rec_start = p;
        if (p + 1 == end) {  // a bare header character at the very end: kseq reports end of file
            if (final) break;
            return done(size_t(rec_start - data));
        }
        p = line_end(p);
        if (p < end) ++p;
        const size_t seq_begin = out.bases.size();
        // sequence lines
        while (p < end && *p != '>' && *p != '@' && *p != '+') {
            if (*p == '\n') {
                ++p;
                continue;
            }
            const char* e = line_end(p);
            size_t len = size_t(e - p);
            // kseq drops the '\r' of "\r\n", but only once the accumulated sequence is longer than one
            // character (ks_getuntil2: `str->l > 1`): a lone "\r" first line stays a base
            if (len && p[len - 1] == '\r' && (out.bases.size() - seq_begin) + len > 1) --len;
            out.bases.insert(out.bases.end(), p, p + len);
            p = e < end ? e + 1 : end;
        }
        const size_t seq_len = out.bases.size() - seq_begin;
        out.offsets.push_back(out.bases.size());
        ++added;
        if (p < end && *p == '+') {  // FASTQ: skip the '+' line, then quality up to the sequence length
            p = line_end(p);
            bool ok = p < end;  // kseq: no newline after '+' = "no quality string" (-2)
            if (p < end) ++p;
            size_t qual = 0;
            // at least one quality line is read, then more while it is shorter than the sequence
            while (ok && p < end) {
                const char* e = line_end(p);
                size_t len = size_t(e - p);
                if (len && p[len - 1] == '\r' && qual + len > 1) --len;  // same rule on the quality string
                qual += len;
                p = e < end ? e + 1 : end;
                if (qual >= seq_len) break;
            }
            if (!ok || qual != seq_len) {
                // kseq_read returns -2 here and the reference's `while (kseq_read(seq) >= 0)` loop ends:
                // the malformed record and everything after it are not processed
                out.bases.resize(seq_begin);
                out.offsets.pop_back();
                --added;
                if (!final && p == end) return done(size_t(rec_start - data));  // perhaps only cut by the chunk
                if (stopped) *stopped = true;
                return done(n);
            }
            while (p < end && *p != '>' && *p != '@') ++p;  // to the next header character
        }
        last_seq_begin = seq_begin;
        last_rec_start = rec_start;
    }
    if (!final && last_rec_start) {  // hold the last record back: it may go on in the next chunk
        out.bases.resize(last_seq_begin);
        out.offsets.pop_back();
        --added;
        return done(size_t(last_rec_start - data));
    }
    return done(n);
}

// Appends every record of `data[0..n)` (a whole file) to `out`.  Returns the number of records appended.
inline uint64_t parse(const char* data, size_t n, Batch& out) {
    uint64_t added = 0;
    parse_some(data, n, out, true, nullptr, &added);
    return added;
}

// Whole file (plain text, or gzip when built with -DLPHASH_B200_WITH_ZLIB -lz) -> records appended
// to `out`.  Throws std::runtime_error if the file cannot be read.
inline uint64_t read_file(std::string const& path, Batch& out) {
    std::vector<char> text;
#ifdef LPHASH_B200_WITH_ZLIB
    gzFile f = gzopen(path.c_str(), "rb");  // transparently reads uncompressed files too
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1u << 20);
    size_t used = 0;
    text.resize(size_t(1) << 24);
    for (;;) {
        if (used == text.size()) text.resize(text.size() * 2);
        const size_t want = text.size() - used;
        int got = gzread(f, text.data() + used, unsigned(want > (1u << 30) ? (1u << 30) : want));
        if (got < 0) {
            gzclose(f);
            throw std::runtime_error("read error in " + path);
        }
        if (got == 0) break;
        used += size_t(got);
    }
    gzclose(f);
    text.resize(used);
#else
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    text.resize(sz > 0 ? size_t(sz) : 0);
    size_t got = text.empty() ? 0 : std::fread(text.data(), 1, text.size(), f);
    std::fclose(f);
    if (got != text.size()) throw std::runtime_error("short read in " + path);
    if (text.size() >= 2 && (unsigned char)text[0] == 0x1f && (unsigned char)text[1] == 0x8b)
        throw std::runtime_error(path + " is gzip-compressed: build with -DLPHASH_B200_WITH_ZLIB -lz");
#endif
    return parse(text.data(), text.size(), out);
}

// ---- streaming ingest --------------------------------------------------------------------------------------
// The file is inflated / read in chunks of `chunk_bytes` on ONE background thread, which also splits the
// records (parse_some) into batches; `on_batch(Batch&)` runs on the calling thread, in file order, while the
// background thread is already producing the next batch (two batches in flight).  With the GPU call inside
// on_batch, inflate + parse of chunk i+1 overlap the H2D copy, the kernels and the D2H copy of chunk i: this
// is the overlap the reference's timed loop (gz + kseq + hf per record, src/query.cpp:48-56) cannot have.
// Returns the number of records delivered.  zlib inflates one stream on one core (~0.3 GB/s of text); a plain
// text file is bounded by the parser (memchr, a few GB/s).
}  // namespace fastx
}  // namespace lphash_b200

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace lphash_b200 {
namespace fastx {

template <class OnBatch>
uint64_t stream_file(std::string const& path, size_t chunk_bytes, OnBatch&& on_batch) {
    if (chunk_bytes < 16) chunk_bytes = 16;
    struct Slot {
        Batch batch;
        bool full = false;
    } slots[2];
    std::mutex mu;
    std::condition_variable cv;
    bool eof = false;
    std::string error;
    std::thread producer([&]() {
        try {
#ifdef LPHASH_B200_WITH_ZLIB
            gzFile f = gzopen(path.c_str(), "rb");  // transparently reads uncompressed files too
            if (!f) throw std::runtime_error("cannot open " + path);
            gzbuffer(f, 1u << 20);
            auto read_some = [&](char* dst, size_t want) -> size_t {
                int got = gzread(f, dst, unsigned(want > (1u << 30) ? (1u << 30) : want));
                if (got < 0) throw std::runtime_error("read error in " + path);
                return size_t(got);
            };
#else
            std::FILE* f = std::fopen(path.c_str(), "rb");
            if (!f) throw std::runtime_error("cannot open " + path);
            auto read_some = [&](char* dst, size_t want) -> size_t { return std::fread(dst, 1, want, f); };
#endif
            std::vector<char> text;  // carry (the record held back) + the next chunk
            size_t carry = 0;
            int which = 0;
            bool more = true, stopped = false;
            while (more && !stopped) {
                text.resize(carry + chunk_bytes);
                size_t got = 0;
                while (got < chunk_bytes) {  // fill the chunk (gzread may return short counts)
                    size_t r = read_some(text.data() + carry + got, chunk_bytes - got);
                    if (r == 0) {
                        more = false;
                        break;
                    }
                    got += r;
                }
#ifndef LPHASH_B200_WITH_ZLIB
                if (carry == 0 && got >= 2 && (unsigned char)text[0] == 0x1f && (unsigned char)text[1] == 0x8b)
                    throw std::runtime_error(path + " is gzip-compressed: build with -DLPHASH_B200_WITH_ZLIB -lz");
#endif
                const size_t n = carry + got;
                Slot& s = slots[which];
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !s.full; });
                }
                s.batch.bases.clear();
                s.batch.offsets.clear();
                const size_t used = parse_some(text.data(), n, s.batch, !more, &stopped);
                carry = n - used;
                if (carry && used) std::memmove(text.data(), text.data() + used, carry);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    s.full = true;
                }
                cv.notify_all();
                which ^= 1;
            }
#ifdef LPHASH_B200_WITH_ZLIB
            gzclose(f);
#else
            std::fclose(f);
#endif
        } catch (std::exception const& e) {
            std::lock_guard<std::mutex> lk(mu);
            error = e.what();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            eof = true;
        }
        cv.notify_all();
    });
    uint64_t delivered = 0;
    int which = 0;
    for (;;) {
        Slot& s = slots[which];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return s.full || eof; });
            if (!s.full) break;  // producer finished and this slot was never filled
        }
        try {
            if (s.batch.n_records()) on_batch(s.batch);
        } catch (...) {
            {
                std::lock_guard<std::mutex> lk(mu);
                s.full = false;
                slots[which ^ 1].full = false;
            }
            cv.notify_all();
            producer.join();  // (it stops at its next wait: both slots are free; the file is read to the end)
            throw;
        }
        delivered += s.batch.n_records();
        {
            std::lock_guard<std::mutex> lk(mu);
            s.full = false;
        }
        cv.notify_all();
        which ^= 1;
    }
    producer.join();
    if (!error.empty()) throw std::runtime_error(error);
    return delivered;
}

}  // namespace fastx
}  // namespace lphash_b200
