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
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// ---- where the text comes from ----------------------------------------------------------------------------------
// Plain file, gzip (one zlib stream on the calling thread, ~0.3 GB/s of text), or BGZF - the blocked gzip `bgzip`
// writes (htslib; every block is a gzip member of at most 64 KB carrying its own size in a "BC" extra field):
// there the blocks of a group are inflated side by side on `threads` threads (LPHB_INGEST_THREADS, default 8),
// which is what lifts compressed input to the parser's speed.  read() fills at most `want` bytes, 0 = end.
class TextSource {
public:
    explicit TextSource(std::string const& path) : path_(path) {
#ifdef LPHASH_B200_WITH_ZLIB
        raw_ = std::fopen(path.c_str(), "rb");
        if (!raw_) throw std::runtime_error("cannot open " + path);
        unsigned char h[18];
        const size_t got = std::fread(h, 1, sizeof h, raw_);
        std::rewind(raw_);
        bgzf_ = got == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C' &&
                h[14] == 2 && h[15] == 0;
        if (const char* e = std::getenv("LPHB_INGEST_THREADS")) {
            long v = std::strtol(e, nullptr, 10);
            if (v >= 1) threads_ = unsigned(v > 64 ? 64 : v);
        }
        if (!bgzf_) {
            std::fclose(raw_);
            raw_ = nullptr;
            gz_ = gzopen(path.c_str(), "rb");  // transparently reads uncompressed files too
            if (!gz_) throw std::runtime_error("cannot open " + path);
            gzbuffer(gz_, 1u << 20);
        }
#else
        raw_ = std::fopen(path.c_str(), "rb");
        if (!raw_) throw std::runtime_error("cannot open " + path);
#endif
    }
    TextSource(TextSource const&) = delete;
    TextSource& operator=(TextSource const&) = delete;
    ~TextSource() {
#ifdef LPHASH_B200_WITH_ZLIB
        if (gz_) gzclose(gz_);
#endif
        if (raw_) std::fclose(raw_);
    }
    bool is_bgzf() const { return bgzf_; }

    size_t read(char* dst, size_t want) {
#ifdef LPHASH_B200_WITH_ZLIB
        if (bgzf_) {
            size_t done = 0;
            while (done < want) {
                if (out_pos_ == out_.size() && !refill()) break;
                const size_t n = std::min(want - done, out_.size() - out_pos_);
                std::memcpy(dst + done, out_.data() + out_pos_, n);
                out_pos_ += n;
                done += n;
            }
            return done;
        }
        int got = gzread(gz_, dst, unsigned(want > (1u << 30) ? (1u << 30) : want));
        if (got < 0) throw std::runtime_error("read error in " + path_);
        return size_t(got);
#else
        const size_t got = std::fread(dst, 1, want, raw_);
        if (first_) {
            first_ = false;
            if (got >= 2 && (unsigned char)dst[0] == 0x1f && (unsigned char)dst[1] == 0x8b)
                throw std::runtime_error(path_ + " is gzip-compressed: build with -DLPHASH_B200_WITH_ZLIB -lz");
        }
        return got;
#endif
    }

private:
#ifdef LPHASH_B200_WITH_ZLIB
    struct Block {
        size_t at, clen, out_at;  // deflate data inside comp_, its length, where the text goes
        uint32_t isize, crc;
    };
    // the next group of blocks (about 32 MB of text), inflated in parallel into out_
    bool refill();
    gzFile gz_ = nullptr;
    std::vector<unsigned char> comp_;
    std::vector<char> out_;
    size_t out_pos_ = 0;
    bool eof_ = false;
#endif
    std::string path_;
    std::FILE* raw_ = nullptr;
    bool bgzf_ = false, first_ = true;
    unsigned threads_ = 8;
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
    TextSource src(path);
    size_t used = 0;
    text.resize(size_t(1) << 24);
    for (;;) {
        if (used == text.size()) text.resize(text.size() * 2);
        const size_t got = src.read(text.data() + used, text.size() - used);
        if (got == 0) break;
        used += got;
    }
    text.resize(used);
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

#ifdef LPHASH_B200_WITH_ZLIB
inline bool TextSource::refill() {
    out_.clear();
    out_pos_ = 0;
    if (eof_) return false;
    comp_.clear();
    std::vector<Block> blocks;
    size_t text_bytes = 0;
    while (text_bytes < (size_t(32) << 20)) {
        unsigned char h[12];
        const size_t got = std::fread(h, 1, sizeof h, raw_);
        if (got == 0) {
            eof_ = true;
            break;
        }
        if (got != sizeof h || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4))
            throw std::runtime_error("bad BGZF block header in " + path_);
        const size_t xlen = size_t(h[10]) | (size_t(h[11]) << 8);
        std::vector<unsigned char> extra(xlen);
        if (std::fread(extra.data(), 1, xlen, raw_) != xlen) throw std::runtime_error("truncated BGZF block in " + path_);
        size_t bsize = 0;
        for (size_t i = 0; i + 4 <= xlen;) {  // subfields: SI1 SI2 SLEN(2) data
            const size_t slen = size_t(extra[i + 2]) | (size_t(extra[i + 3]) << 8);
            if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen)
                bsize = (size_t(extra[i + 4]) | (size_t(extra[i + 5]) << 8)) + 1;
            i += 4 + slen;
        }
        if (bsize < 12 + xlen + 8) throw std::runtime_error("BGZF block without a valid BC field in " + path_);
        const size_t rest = bsize - 12 - xlen;  // deflate data + CRC32 + ISIZE
        const size_t at = comp_.size();
        comp_.resize(at + rest);
        if (std::fread(comp_.data() + at, 1, rest, raw_) != rest) throw std::runtime_error("truncated BGZF block in " + path_);
        const unsigned char* t = comp_.data() + at + rest - 8;
        Block b;
        b.at = at;
        b.clen = rest - 8;
        b.crc = uint32_t(t[0]) | (uint32_t(t[1]) << 8) | (uint32_t(t[2]) << 16) | (uint32_t(t[3]) << 24);
        b.isize = uint32_t(t[4]) | (uint32_t(t[5]) << 8) | (uint32_t(t[6]) << 16) | (uint32_t(t[7]) << 24);
        b.out_at = text_bytes;
        text_bytes += b.isize;
        if (b.isize) blocks.push_back(b);  // (the empty end-of-file marker block carries nothing)
    }
    if (blocks.empty()) return !eof_ ? refill() : false;
    out_.resize(text_bytes);
    const unsigned nt = unsigned(std::min<size_t>(threads_, blocks.size()));
    std::vector<std::string> errs(nt);
    auto work = [&](unsigned t) {
        z_stream zs;
        std::memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) {
            errs[t] = "inflateInit2 failed";
            return;
        }
        for (size_t i = t; i < blocks.size(); i += nt) {
            Block const& b = blocks[i];
            inflateReset(&zs);
            zs.next_in = comp_.data() + b.at;
            zs.avail_in = uInt(b.clen);
            zs.next_out = reinterpret_cast<Bytef*>(out_.data() + b.out_at);
            zs.avail_out = b.isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0 ||
                uint32_t(crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef*>(out_.data() + b.out_at), b.isize)) != b.crc) {
                errs[t] = "corrupt BGZF block in " + path_;
                break;
            }
        }
        inflateEnd(&zs);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (auto const& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
    return true;
}
#endif

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
            TextSource src(path);
            auto read_some = [&](char* dst, size_t want) -> size_t { return src.read(dst, want); };
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
