// bam_input.hpp — SAM text source for the sam2pairs executable that also accepts BAM (second half: "BAM
// (BGZF) input decoding to skip `samtools view`", microcket:478,500).  Host code only.  SamSource::read() is a drop-in for
// fread(): plain SAM text is passed through byte for byte; a stream that starts with a BGZF block holding the BAM magic is
// decoded to the text `samtools view` (no -h) prints, which is what the reference's sam2pairs sees on /dev/stdin.
//   reader thread -> batches of BGZF blocks -> inflated side by side by worker threads (one raw-deflate stream each, CRC-32 and
//   ISIZE checked) -> alignment records (they may straddle blocks) -> text, again in parallel slices -> bounded queue -> read().
// The device path is unchanged: it still tokenises text.  A BAM-native device parser (fixed-offset fields instead of text
// tokenising) is the next step and is NOT built.
// BAM stores bases as 4-bit codes: lower case and characters such as '.' do not survive the aligner's SAM -> BAM step (they come
// back upper case / as N, exactly as through `samtools view`), which matters only to cfg.rmdup's treatment of soft-masked reads.
// Parity: the decoder follows the SAM/BAM specification (SAMv1 §4.2); samtools is not in this image, so the tests encode BAM
// with an independent Python writer (tests/bam_writer.py) and require the decoded text to equal the SAM it was made from.
#pragma once
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mkbam {

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

// text goes through a raw pointer into a buffer sized by text_bound(): no per-character capacity checks
static inline char *put_u(char *o, uint64_t v) { char b[24]; int n = 0; do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) *o++ = b[--n]; return o; }
static inline char *put_i(char *o, int64_t v) { if (v < 0) { *o++ = '-'; return put_u(o, (uint64_t)(-(v + 1)) + 1); } return put_u(o, (uint64_t)v); }
static inline char *put_s(char *o, const std::string &s) { memcpy(o, s.data(), s.size()); return o + s.size(); }

// upper bound of the text of a record of `len` bytes: the densest field is an int8 array element (1 byte -> ",-128"); the 32
// fixed bytes become at most five numbers, two reference names and eleven tabs
static inline size_t text_bound(size_t len, size_t max_ref) { return 5 * len + 2 * max_ref + 96; }

struct NibblePairs { char t[256][2]; NibblePairs() { const char *a = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 256; ++i) { t[i][0] = a[i >> 4]; t[i][1] = a[i & 15]; } } };

// one alignment record (without its block_size word) -> one SAM line at o; returns the end of the line, NULL when the
// record does not fit its own length
static inline char *rec_to_sam(const uint8_t *r, size_t len, const std::vector<std::string> &refs, char *o) {
    static const NibblePairs NP;
    if (len < 32) return NULL;
    const int32_t ref = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
    const uint32_t l_name = r[8], mapq = r[9], n_cig = rd16(r + 12), flag = rd16(r + 14), l_seq = rd32(r + 16);
    const int32_t nref = (int32_t)rd32(r + 20), npos = (int32_t)rd32(r + 24), tlen = (int32_t)rd32(r + 28);
    size_t p = 32;
    const size_t need = (size_t)l_name + (size_t)n_cig * 4 + ((size_t)l_seq + 1) / 2 + l_seq;
    if (l_name == 0 || p + need > len || ref >= (int32_t)refs.size() || nref >= (int32_t)refs.size()) return NULL;
    memcpy(o, r + p, l_name - 1); o += l_name - 1; p += l_name;
    *o++ = '\t'; o = put_u(o, flag);
    *o++ = '\t'; if (ref < 0) *o++ = '*'; else o = put_s(o, refs[ref]);
    *o++ = '\t'; o = put_i(o, (int64_t)pos + 1);
    *o++ = '\t'; o = put_u(o, mapq);
    *o++ = '\t';
    if (n_cig == 0) *o++ = '*';
    for (uint32_t k = 0; k < n_cig; ++k, p += 4) { const uint32_t c = rd32(r + p); o = put_u(o, c >> 4); *o++ = "MIDNSHP=X???????"[c & 15]; }
    *o++ = '\t'; if (nref < 0) *o++ = '*'; else if (nref == ref) *o++ = '='; else o = put_s(o, refs[nref]);
    *o++ = '\t'; o = put_i(o, (int64_t)npos + 1);
    *o++ = '\t'; o = put_i(o, tlen);
    *o++ = '\t';
    if (l_seq == 0) *o++ = '*';
    for (uint32_t k = 0; k + 1 < l_seq; k += 2) { const char *t = NP.t[r[p + (k >> 1)]]; o[k] = t[0]; o[k + 1] = t[1]; }
    if (l_seq & 1) o[l_seq - 1] = NP.t[r[p + (l_seq >> 1)]][0];
    o += l_seq; p += ((size_t)l_seq + 1) / 2;
    *o++ = '\t';
    if (l_seq == 0 || r[p] == 0xFF) *o++ = '*'; else { for (uint32_t k = 0; k < l_seq; ++k) o[k] = (char)(r[p + k] + 33); o += l_seq; }
    p += l_seq;
    while (p < len) {                                                   // optional fields: tag[2] type value
        if (p + 3 > len) return NULL;
        *o++ = '\t'; *o++ = (char)r[p]; *o++ = (char)r[p + 1]; *o++ = ':';
        const char t = (char)r[p + 2]; p += 3;
        auto scalar = [&](char ty) -> bool {
            switch (ty) {
                case 'c': if (p + 1 > len) return false; o = put_i(o, (int8_t)r[p]); p += 1; return true;
                case 'C': if (p + 1 > len) return false; o = put_u(o, r[p]); p += 1; return true;
                case 's': if (p + 2 > len) return false; o = put_i(o, (int16_t)rd16(r + p)); p += 2; return true;
                case 'S': if (p + 2 > len) return false; o = put_u(o, rd16(r + p)); p += 2; return true;
                case 'i': if (p + 4 > len) return false; o = put_i(o, (int32_t)rd32(r + p)); p += 4; return true;
                case 'I': if (p + 4 > len) return false; o = put_u(o, rd32(r + p)); p += 4; return true;
                case 'f': { if (p + 4 > len) return false; float f; memcpy(&f, r + p, 4); o += snprintf(o, 16, "%g", f); p += 4; return true; }
                default: return false;
            }
        };
        if (t == 'A') { if (p + 1 > len) return NULL; *o++ = 'A'; *o++ = ':'; *o++ = (char)r[p++]; }
        else if (t == 'Z' || t == 'H') {
            *o++ = t; *o++ = ':';
            const void *e = memchr(r + p, 0, len - p); if (!e) return NULL;
            const size_t n = (const uint8_t *)e - (r + p); memcpy(o, r + p, n); o += n; p += n + 1;
        } else if (t == 'B') {
            if (p + 5 > len) return NULL;
            const char st = (char)r[p]; const uint32_t cnt = rd32(r + p + 1); p += 5;
            *o++ = 'B'; *o++ = ':'; *o++ = st;
            for (uint32_t k = 0; k < cnt; ++k) { *o++ = ','; if (!scalar(st)) return NULL; }
        } else if (t == 'f') { *o++ = 'f'; *o++ = ':'; if (!scalar('f')) return NULL; }
        else { *o++ = 'i'; *o++ = ':'; if (!scalar(t)) return NULL; }  // every integer width prints as type i
    }
    *o++ = '\n';
    return o;
}

struct Chunk { std::unique_ptr<char[]> p; size_t n = 0; };              // a slice of decoded text

class SamSource {
public:
    explicit SamSource(FILE *f) : f_(f) {
        npeek_ = fread(peek_, 1, sizeof peek_, f_);
        bam_ = npeek_ >= 18 && peek_[0] == 0x1f && peek_[1] == 0x8b && peek_[2] == 8 && (peek_[3] & 4);
        if (bam_) {
            nthreads_ = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
            if (const char *e = getenv("MICROCKET_BAM_THREADS")) nthreads_ = (unsigned)std::max(1, atoi(e));
            producer_ = std::thread([this] { produce(); });
        }
    }
    ~SamSource() {
        if (producer_.joinable()) {                                    // normally the stream has ended and the thread is done
            std::unique_lock<std::mutex> l(m_);
            stop_ = true; cv_.notify_all();
            const bool ended = cv_.wait_for(l, std::chrono::milliseconds(200), [this] { return done_; });
            l.unlock();
            if (ended) producer_.join(); else producer_.detach();      // error exit while it still waits on a pipe: the process is about to end
        }
    }
    bool is_bam() const { return bam_; }
    bool failed() const { return failed_.load(); }
    const std::string &error() const { return err_; }

    // same contract as fread(dst, 1, n, f): short only at the end of the stream (or after an error: failed())
    size_t read(char *dst, size_t n) {
        if (!bam_) {
            size_t got = 0;
            if (peek_off_ < npeek_) { got = std::min(n, npeek_ - peek_off_); memcpy(dst, peek_ + peek_off_, got); peek_off_ += got; }
            if (got < n) got += fread(dst + got, 1, n - got, f_);
            return got;
        }
        size_t got = 0;
        while (got < n) {
            if (cur_off_ == cur_.n) {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return !q_.empty() || done_; });
                if (q_.empty()) break;
                cur_ = std::move(q_.front()); q_.pop_front(); cur_off_ = 0;
                l.unlock(); cv_.notify_all();
                continue;
            }
            const size_t m = std::min(n - got, cur_.n - cur_off_);
            memcpy(dst + got, cur_.p.get() + cur_off_, m); got += m; cur_off_ += m;
        }
        return got;
    }

private:
    struct Block { size_t in_off, in_len, out_off, out_len; };

    size_t raw_read(uint8_t *dst, size_t n) {                          // the peeked bytes first
        size_t got = 0;
        if (peek_off_ < npeek_) { got = std::min(n, npeek_ - peek_off_); memcpy(dst, peek_ + peek_off_, got); peek_off_ += got; }
        if (got < n) got += fread(dst + got, 1, n - got, f_);
        return got;
    }
    void fail(const std::string &why) { err_ = why; failed_ = true; }
    void push(Chunk &&s) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return q_.size() < 256 || stop_; });
        if (!stop_) q_.push_back(std::move(s));
        l.unlock(); cv_.notify_all();
    }
    template <class F> void parallel(size_t n, F fn) {                  // fn(i) for i in [0, n), nthreads_ at a time
        std::atomic<size_t> next(0);
        auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthreads_ && t < n; ++t) th.emplace_back(work);
        work(); for (auto &t : th) t.join();
    }

    void produce() {
        const size_t BATCH = 512;                                      // BGZF blocks per round (<= 32 MiB inflated)
        std::vector<uint8_t> in, raw;                                  // raw: carried-over partial record + this round's bytes
        std::vector<Block> blocks;
        size_t carry = 0; bool eof = false, header_done = false;
        size_t hdr_need = 0; int hdr_state = 0; int32_t n_ref = 0, refs_seen = 0;
        while (!eof && !failed_) {
            { std::lock_guard<std::mutex> l(m_); if (stop_) break; }
            in.clear(); blocks.clear();
            size_t out_total = 0;
            while (blocks.size() < BATCH) {
                uint8_t h[18];
                const size_t g = raw_read(h, 18);
                if (g == 0) { eof = true; break; }
                if (g < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { fail("BAM input: not a BGZF block"); break; }
                const size_t xlen = rd16(h + 10);
                size_t off = in.size(); in.resize(off + 18); memcpy(in.data() + off, h, 18);
                if (xlen > 6) { in.resize(off + 12 + xlen); if (raw_read(in.data() + off + 18, xlen - 6) != xlen - 6) { fail("BAM input: truncated BGZF header"); break; } }
                int bsize = -1;                                        // the BC subfield may sit anywhere in the extra field
                for (size_t x = 0; x + 4 <= xlen;) {
                    const uint8_t *s = in.data() + off + 12 + x; const size_t sl = rd16(s + 2);
                    if (s[0] == 'B' && s[1] == 'C' && sl == 2 && x + 6 <= xlen) bsize = rd16(s + 4);
                    x += 4 + sl;
                }
                if (bsize < 0 || (size_t)bsize + 1 < 12 + xlen + 8) { fail("BAM input: BGZF block without a BC field"); break; }
                const size_t rest = (size_t)bsize + 1 - 12 - xlen;     // deflate data + CRC32 + ISIZE
                const size_t d0 = in.size(); in.resize(d0 + rest);
                if (raw_read(in.data() + d0, rest) != rest) { fail("BAM input: truncated BGZF block"); break; }
                const size_t isize = rd32(in.data() + d0 + rest - 4);
                if (isize > 65536) { fail("BAM input: BGZF block larger than 64 KiB"); break; }
                blocks.push_back({d0, rest, out_total, isize}); out_total += isize;
            }
            if (failed_) break;
            raw.resize(carry + out_total);
            std::atomic<bool> bad(false);
            parallel(blocks.size(), [&](size_t i) {
                const Block &b = blocks[i];
                if (b.out_len == 0) return;
                z_stream zs; memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
                zs.next_in = in.data() + b.in_off; zs.avail_in = (uInt)(b.in_len - 8);
                zs.next_out = raw.data() + carry + b.out_off; zs.avail_out = (uInt)b.out_len;
                const int rc = inflate(&zs, Z_FINISH);
                const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
                inflateEnd(&zs);
                if (!ok || crc32(crc32(0L, Z_NULL, 0), raw.data() + carry + b.out_off, (uInt)b.out_len) != rd32(in.data() + b.in_off + b.in_len - 8)) bad = true;
            });
            if (bad) { fail("BAM input: a BGZF block does not inflate to its recorded size / CRC"); break; }
            size_t p = 0; const size_t end = raw.size();
            if (!header_done) {                                        // magic, header text, reference names: may span many blocks
                while (true) {
                    if (hdr_state == 0) { if (end - p < 8) break; if (memcmp(raw.data() + p, "BAM\1", 4)) { fail("BAM input: BGZF data without the BAM magic"); break; } hdr_need = rd32(raw.data() + p + 4); p += 8; hdr_state = 1; }
                    else if (hdr_state == 1) { if (end - p < hdr_need + 4) break; p += hdr_need; n_ref = (int32_t)rd32(raw.data() + p); p += 4; hdr_state = 2; }
                    else if (refs_seen < n_ref) {
                        if (end - p < 4) break;
                        const size_t ln = rd32(raw.data() + p);
                        if (end - p < 4 + ln + 4) break;
                        refs_.emplace_back((const char *)raw.data() + p + 4, ln ? ln - 1 : 0); p += 8 + ln; ++refs_seen;
                        max_ref_ = std::max(max_ref_, refs_.back().size());
                    } else { header_done = true; break; }
                }
                if (failed_) break;
            }
            std::vector<std::pair<size_t, size_t>> recs;               // (offset behind block_size, length)
            if (header_done)
                while (end - p >= 4) {
                    const size_t bl = rd32(raw.data() + p);
                    if (bl > (1u << 28)) { fail("BAM input: implausible record length"); break; }
                    if (end - p < 4 + bl) break;
                    recs.push_back({p + 4, bl}); p += 4 + bl;
                }
            if (failed_) break;
            if (!recs.empty()) {
                const size_t slices = std::min<size_t>(recs.size(), (size_t)nthreads_ * 4);
                std::vector<Chunk> part(slices);
                parallel(slices, [&](size_t s) {
                    const size_t a = recs.size() * s / slices, b = recs.size() * (s + 1) / slices;
                    const size_t bytes = recs[b - 1].first + recs[b - 1].second - recs[a].first;
                    Chunk &c = part[s];
                    c.p.reset(new char[text_bound(bytes, max_ref_) + (b - a) * (2 * max_ref_ + 96)]);   // untouched pages cost nothing
                    char *o = c.p.get();
                    for (size_t k = a; k < b; ++k) { o = rec_to_sam(raw.data() + recs[k].first, recs[k].second, refs_, o); if (!o) { bad = true; return; } }
                    c.n = (size_t)(o - c.p.get());
                });
                if (bad) { fail("BAM input: malformed alignment record"); break; }
                for (auto &s : part) push(std::move(s));               // read() takes the slices in order: no concatenation
            }
            carry = end - p;
            if (carry && p) memmove(raw.data(), raw.data() + p, carry);
            raw.resize(carry);
        }
        if (!failed_ && (carry || !header_done)) fail("BAM input: truncated file");
        { std::lock_guard<std::mutex> l(m_); done_ = true; }
        cv_.notify_all();
    }

    FILE *f_; uint8_t peek_[18]; size_t npeek_ = 0, peek_off_ = 0; bool bam_ = false;
    unsigned nthreads_ = 1;
    std::vector<std::string> refs_; size_t max_ref_ = 0;
    std::thread producer_; std::mutex m_; std::condition_variable cv_;
    std::deque<Chunk> q_; bool done_ = false, stop_ = false;
    Chunk cur_; size_t cur_off_ = 0;
    std::atomic<bool> failed_{false}; std::string err_;
};

}  // namespace mkbam
