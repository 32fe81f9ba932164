// host_bam.cpp -- see host_bam.h
#include "host_bam.h"
#include "bsb_bam.h"
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#include <algorithm>
#include <chrono>
#include <stdexcept>
#include <thread>

namespace bsb {

namespace {
constexpr size_t BGZF_IN = 0xff00;      // uncompressed bytes per block (htslib BGZF_BLOCK_SIZE)
constexpr size_t BGZF_HDR = 18, BGZF_TAIL = 8;
const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline void put32(std::vector<uint8_t> &o, uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; o.insert(o.end(), b, b + 4); }
inline void set32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

[[noreturn]] void bad(const char *what, const char *p, const char *e)
{
    throw std::runtime_error(std::string("[E::bam_encode] ") + what + ": " + std::string(p, std::min<size_t>(e - p, 80)));
}

} // namespace

struct BamWriter::Worker {
    z_stream zs;
    bool z_ok = false;
    std::vector<uint8_t> raw, comp;
    uint64_t n_rec = 0, n_raw = 0;
    std::string err;

    void init(int level)
    {
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("[E::bam] deflateInit2 failed");
        z_ok = true;
        raw.reserve(BGZF_IN + 4096);
    }
    ~Worker() { if (z_ok) deflateEnd(&zs); }

    // one BGZF block from raw[0, n)
    void block(const uint8_t *src, size_t n)
    {
        const size_t at = comp.size();
        comp.resize(at + 0x10000);
        uint8_t *dst = comp.data() + at;
        static const uint8_t H[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
        memcpy(dst, H, 16);
        deflateReset(&zs);
        zs.next_in = const_cast<Bytef *>(src); zs.avail_in = (uInt)n;
        zs.next_out = dst + BGZF_HDR; zs.avail_out = (uInt)(0x10000 - BGZF_HDR - BGZF_TAIL);
        size_t clen;
        if (deflate(&zs, Z_FINISH) == Z_STREAM_END) clen = zs.total_out;
        else {   // did not fit (incompressible input): one stored deflate block, 5 bytes of overhead
            uint8_t *d = dst + BGZF_HDR;
            d[0] = 1; d[1] = (uint8_t)n; d[2] = (uint8_t)(n >> 8); d[3] = (uint8_t)~d[1]; d[4] = (uint8_t)~d[2];
            memcpy(d + 5, src, n);
            clen = n + 5;
        }
        const size_t total = BGZF_HDR + clen + BGZF_TAIL;
        dst[16] = (uint8_t)(total - 1); dst[17] = (uint8_t)((total - 1) >> 8);
        set32(dst + BGZF_HDR + clen, (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n));
        set32(dst + BGZF_HDR + clen + 4, (uint32_t)n);
        comp.resize(at + total);
    }
    void flush_raw()
    {
        size_t off = 0;
        while (off < raw.size()) { const size_t n = std::min(BGZF_IN, raw.size() - off); block(raw.data() + off, n); off += n; }
        n_raw += raw.size();
        raw.clear();
    }
};

BamWriter::BamWriter(const std::string &path, int threads, int level) : threads_(std::max(1, threads)), level_(level)
{
    if (level_ < -1 || level_ > 9) level_ = -1;
    f_ = fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("[E::bam] cannot open " + path + " for writing: " + strerror(errno));
    setvbuf(f_, nullptr, _IOFBF, 1 << 22);
    for (int t = 0; t < threads_; ++t) { workers_.push_back(new Worker); workers_.back()->init(level_); }
}

BamWriter::~BamWriter()
{
    try { close(); } catch (...) {}
    for (Worker *w : workers_) delete w;
}

int BamWriter::tid_of(const char *p, size_t n) const
{
    auto it = ref_ids_.find(std::string(p, n));
    return it == ref_ids_.end() ? -1 : it->second;
}

void BamWriter::header(const std::string &text)
{
    if (have_header_) throw std::runtime_error("[E::bam] header written twice");
    have_header_ = true;
    std::vector<std::pair<std::string, uint32_t>> refs;
    size_t p = 0;
    while (p < text.size()) {
        size_t e = text.find('\n', p);
        if (e == std::string::npos) e = text.size();
        if (e - p >= 3 && text.compare(p, 3, "@SQ") == 0) {
            std::string name; uint64_t len = 0;
            size_t q = p + 3;
            while (q < e) {
                if (text[q] == '\t') { ++q; continue; }
                size_t t = text.find('\t', q);
                if (t == std::string::npos || t > e) t = e;
                if (t - q > 3 && text[q + 2] == ':') {
                    if (text.compare(q, 3, "SN:") == 0) name = text.substr(q + 3, t - q - 3);
                    else if (text.compare(q, 3, "LN:") == 0) len = strtoull(text.c_str() + q + 3, nullptr, 10);
                }
                q = t;
            }
            if (!name.empty() && !ref_ids_.count(name)) { ref_ids_[name] = (int)refs.size(); refs.emplace_back(name, (uint32_t)len); }
        }
        p = e + 1;
    }
    Worker &w = *workers_[0];
    std::vector<uint8_t> &o = w.raw;
    o.insert(o.end(), {'B', 'A', 'M', 1});
    put32(o, (uint32_t)text.size());
    o.insert(o.end(), text.begin(), text.end());
    put32(o, (uint32_t)refs.size());
    for (auto &r : refs) {
        put32(o, (uint32_t)r.first.size() + 1);
        o.insert(o.end(), r.first.begin(), r.first.end());
        o.push_back(0);
        put32(o, r.second);
    }
    w.flush_raw();
    raw_bytes_ += w.n_raw; w.n_raw = 0;
    if (fwrite(w.comp.data(), 1, w.comp.size(), f_) != w.comp.size()) throw std::runtime_error("[E::bam] writing the BAM header failed");
    file_bytes_ += w.comp.size();
    w.comp.clear();
}

// sam_parse1 (sam.c:1924-2160) followed by bam_write1 (sam.c:661-735): bam_record (bsb_bam.h), the statements the device runs too
void BamWriter::encode_record(const char *p, const char *e, std::vector<uint8_t> &o) const
{
    auto tid = [this](const char *s, size_t n) { return tid_of(s, n); };
    SamCount c;
    BamCore core;
    int rc = bam_record(c, tid, p, e, nullptr, &core);
    if (rc != BAM_OK) bad(bam_strerror(rc), p, e);
    const size_t at = o.size();
    o.resize(at + c.n);
    SamWrite w = {reinterpret_cast<char *>(o.data() + at)};
    rc = bam_record(w, tid, p, e, &core, nullptr);
    if (rc != BAM_OK || (size_t)(w.p - reinterpret_cast<char *>(o.data() + at)) != c.n) bad("record changed between sizing and writing", p, e);
}

// BGZF blocks made elsewhere (the device's deflate, bsb_deflate.h): they only have to reach the file, in order
void BamWriter::blocks(const uint8_t *bgzf, size_t n, uint64_t raw_bytes, uint64_t n_records)
{
    if (!have_header_) throw std::runtime_error("[E::bam] records before the header");
    const double t0 = now_s();
    if (n && fwrite(bgzf, 1, n, f_) != n) throw std::runtime_error("[E::bam] write failed");
    file_bytes_ += n; raw_bytes_ += raw_bytes; n_records_ += n_records;
    sec_busy_ += now_s() - t0;
}

void BamWriter::records(const char *text, size_t n)
{
    if (!n) return;
    if (!have_header_) throw std::runtime_error("[E::bam] records before the header");
    const double t0 = now_s();
    const char *end = text + n;
    // cut the text into one run of whole lines per worker
    int nt = (int)std::min<size_t>((size_t)threads_, n / 65536 + 1);
    std::vector<const char *> cut(nt + 1);
    cut[0] = text; cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char *p = text + n / nt * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        cut[t] = nl ? nl + 1 : end;
    }
    auto work = [&](int t) {
        Worker &w = *workers_[t];
        try {
            const char *p = cut[t], *pe = cut[t + 1];
            while (p < pe) {
                const char *nl = (const char *)memchr(p, '\n', (size_t)(pe - p));
                const char *le = nl ? nl : pe;
                const char *re = le;
                if (re > p && re[-1] == '\r') --re;
                if (re > p) {
                    const size_t before = w.raw.size();
                    encode_record(p, re, w.raw);
                    ++w.n_rec;
                    if (w.raw.size() > BGZF_IN && before > 0) {   // the record does not fit: close the block in front of it
                        w.block(w.raw.data(), before);
                        w.n_raw += before;
                        w.raw.erase(w.raw.begin(), w.raw.begin() + (long)before);
                    }
                    if (w.raw.size() > BGZF_IN) w.flush_raw();    // a record larger than a block spans several
                }
                p = nl ? nl + 1 : pe;
            }
            w.flush_raw();
        } catch (const std::exception &ex) { w.err = ex.what(); }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < nt; ++t) {
        Worker &w = *workers_[t];
        if (!w.err.empty()) { std::string m = w.err; w.err.clear(); throw std::runtime_error(m); }
        if (fwrite(w.comp.data(), 1, w.comp.size(), f_) != w.comp.size()) throw std::runtime_error("[E::bam] write failed");
        file_bytes_ += w.comp.size(); raw_bytes_ += w.n_raw; n_records_ += w.n_rec;
        w.comp.clear(); w.n_raw = 0; w.n_rec = 0;
    }
    sec_busy_ += now_s() - t0;
}

void BamWriter::close()
{
    if (closed_ || !f_) return;
    closed_ = true;
    if (!have_header_) header("");
    const bool short_write = fwrite(BGZF_EOF, 1, sizeof BGZF_EOF, f_) != sizeof BGZF_EOF || fflush(f_) != 0 || ferror(f_);
    file_bytes_ += sizeof BGZF_EOF;
    const int rc = fclose(f_);
    f_ = nullptr;
    if (short_write || rc) throw std::runtime_error("[E::bam] writing the end of the BAM file failed (disk full?)");
}

uint64_t stream_bam(int in_fd, const std::string &bam_path, int threads, int level)
{
    BamWriter bw(bam_path, threads, level);
    std::string buf, hdr;
    bool in_header = true;
    const size_t CHUNK = 32u << 20;
    size_t have = 0;
    buf.resize(CHUNK);
    for (;;) {
        if (have == buf.size()) buf.resize(buf.size() * 2);          // a line longer than the buffer
        const ssize_t r = read(in_fd, &buf[have], buf.size() - have);
        if (r < 0) { if (errno == EINTR) continue; throw std::runtime_error(std::string("[E::stream_bam] read failed: ") + strerror(errno)); }
        const bool eof = r == 0;
        have += (size_t)r;
        size_t upto = have;
        if (!eof) {
            const void *nl = memrchr(buf.data(), '\n', have);
            if (!nl) continue;
            upto = (size_t)((const char *)nl - buf.data()) + 1;
        }
        size_t p = 0;
        while (in_header && p < upto) {
            if (buf[p] != '@') { in_header = false; bw.header(hdr); break; }
            const void *nl = memchr(buf.data() + p, '\n', upto - p);
            const size_t le = nl ? (size_t)((const char *)nl - buf.data()) + 1 : upto;
            hdr.append(buf, p, le - p);
            if (!nl) hdr.push_back('\n');
            p = le;
        }
        if (!in_header && p < upto) bw.records(buf.data() + p, upto - p);
        memmove(&buf[0], buf.data() + upto, have - upto);
        have -= upto;
        if (eof) break;
    }
    if (in_header) bw.header(hdr);
    bw.close();
    return bw.n_records();
}

} // namespace bsb
