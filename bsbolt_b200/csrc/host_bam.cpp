// host_bam.cpp -- see host_bam.h
#include "host_bam.h"
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

// seq_nt16_table (hts.c:73-91): "=ACMGRSVTWYHKDBN", anything else 15
struct Nt16 {
    uint8_t t[256];
    Nt16()
    {
        memset(t, 15, sizeof t);
        const char *s = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; ++i) { t[(uint8_t)s[i]] = (uint8_t)i; if (s[i] >= 'A') t[(uint8_t)(s[i] + 32)] = (uint8_t)i; }
        t[(uint8_t)'0'] = 1; t[(uint8_t)'1'] = 2; t[(uint8_t)'2'] = 4; t[(uint8_t)'3'] = 8;
    }
};
const Nt16 g_nt16;

inline int cigar_op(char c)
{
    switch (c) {
    case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4;
    case 'H': return 5; case 'P': return 6; case '=': return 7; case 'X': return 8; case 'B': return 9;
    }
    return -1;
}

// hts_reg2bin(beg, end, 14, 5) (htslib/hts.h:1322-1328)
inline int reg2bin(int64_t beg, int64_t end)
{
    int l, s = 14, t = ((1 << 15) - 1) / 7;
    for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << ((l << 1) + l))
        if (beg >> s == end >> s) return t + (int)(beg >> s);
    return 0;
}

inline void put32(std::vector<uint8_t> &o, uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; o.insert(o.end(), b, b + 4); }
inline void set32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

[[noreturn]] void bad(const char *what, const char *p, const char *e)
{
    throw std::runtime_error(std::string("[E::bam_encode] ") + what + ": " + std::string(p, std::min<size_t>(e - p, 80)));
}

inline uint64_t parse_uint(const char *&p, const char *e)
{
    uint64_t v = 0;
    if (p < e && *p == '+') ++p;
    while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return v;
}
inline int64_t parse_int(const char *&p, const char *e)
{
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) neg = *p++ == '-';
    const uint64_t v = parse_uint(p, e);
    return neg ? -(int64_t)v : (int64_t)v;
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

// sam_parse1 (sam.c:1924-2160) followed by bam_write1 (sam.c:661-735), little-endian host
void BamWriter::encode_record(const char *p, const char *e, std::vector<uint8_t> &o) const
{
    const char *line = p;
    auto field = [&](const char *&b, size_t &n) {   // next tab-separated field; p moves behind its tab
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) bad("truncated record", line, e);
        b = p; n = (size_t)(t - p); p = t + 1;
    };
    const char *qn, *s; size_t l_qn, n;
    field(qn, l_qn);
    if (l_qn + 1 > 255) bad("query name too long", line, e);
    uint32_t flag = (uint32_t)parse_uint(p, e);
    if (p >= e || *p++ != '\t') bad("malformed FLAG", line, e);
    field(s, n);
    int32_t tid = (n == 1 && *s == '*') ? -1 : tid_of(s, n);
    int64_t pos = (int64_t)parse_uint(p, e) - 1;
    if (p >= e || *p++ != '\t') bad("malformed POS", line, e);
    if (pos < 0 && tid >= 0) tid = -1;
    if (tid < 0) flag |= 4;
    const uint32_t mapq = (uint32_t)parse_uint(p, e) & 0xff;
    if (p >= e || *p++ != '\t') bad("malformed MAPQ", line, e);

    const size_t at = o.size();
    o.resize(at + 36 + l_qn + 1);
    memcpy(o.data() + at + 36, qn, l_qn);
    o[at + 36 + l_qn] = 0;

    uint32_t n_cigar = 0;
    int64_t cigreflen = 1, qlen_cigar = 0;
    if (*p != '*') {
        int64_t rlen = 0;
        while (p < e && *p != '\t') {
            const uint64_t len = parse_uint(p, e);
            const int op = p < e ? cigar_op(*p) : -1;
            if (op < 0) bad("unrecognized CIGAR operator", line, e);
            ++p;
            put32(o, (uint32_t)(len << 4) | (uint32_t)op);
            ++n_cigar;
            if ((0x3C1A7 >> (op << 1)) & 2) rlen += (int64_t)len;
            if ((0x3C1A7 >> (op << 1)) & 1) qlen_cigar += (int64_t)len;
        }
        if (p >= e || *p++ != '\t') bad("truncated record", line, e);
        if (n_cigar == 0) bad("no CIGAR operations", line, e);
        if (n_cigar > 0xffff) bad("more than 65535 CIGAR operations", line, e);
        cigreflen = !(flag & 4) ? rlen : 1;
    } else {
        flag |= 4;
        field(s, n);
    }
    const int bin = reg2bin(pos, pos + cigreflen);
    field(s, n);
    int32_t mtid;
    if (n == 1 && *s == '=') mtid = tid;
    else if (n == 1 && *s == '*') mtid = -1;
    else mtid = tid_of(s, n);
    int64_t mpos = (int64_t)parse_uint(p, e) - 1;
    if (p >= e || *p++ != '\t') bad("malformed PNEXT", line, e);
    if (mpos < 0 && mtid >= 0) mtid = -1;
    const int64_t isize = parse_int(p, e);
    if (p >= e || *p++ != '\t') bad("malformed TLEN", line, e);
    field(s, n);
    uint32_t l_seq = 0;
    if (!(n == 1 && *s == '*')) {
        l_seq = (uint32_t)n;
        if (n_cigar && qlen_cigar != (int64_t)l_seq) bad("CIGAR and query sequence are of different length", line, e);
        const size_t a = o.size();
        o.resize(a + (l_seq + 1) / 2);
        uint8_t *t = o.data() + a;
        uint32_t i = 0;
        for (; i + 1 < l_seq; i += 2) t[i >> 1] = (uint8_t)(g_nt16.t[(uint8_t)s[i]] << 4 | g_nt16.t[(uint8_t)s[i + 1]]);
        if (i < l_seq) t[i >> 1] = (uint8_t)(g_nt16.t[(uint8_t)s[i]] << 4);
    }
    {   // QUAL: the last mandatory field, ends at a tab or at the end of the line
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *qe = t ? t : e;
        const size_t a = o.size();
        o.resize(a + l_seq);
        if (qe - p == 1 && *p == '*') memset(o.data() + a, 0xff, l_seq);
        else {
            if ((size_t)(qe - p) != l_seq) bad("SEQ and QUAL are of different length", line, e);
            for (uint32_t i = 0; i < l_seq; ++i) {
                const int v = (uint8_t)p[i] - 33;
                if (v < 0 || v > 127) bad("invalid QUAL character", line, e);
                o[a + i] = (uint8_t)v;
            }
        }
        p = t ? t + 1 : e;
    }
    while (p < e) {   // optional fields TAG:TYPE:VALUE
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *fe = t ? t : e;
        if (fe - p < 5 || p[2] != ':' || p[4] != ':') bad("incomplete aux field", line, e);
        const char type = p[3];
        const char *v = p + 5;
        o.push_back((uint8_t)p[0]); o.push_back((uint8_t)p[1]);
        if (type == 'A' || type == 'a' || type == 'c' || type == 'C') {
            if (v >= fe) bad("incomplete aux field", line, e);
            o.push_back('A'); o.push_back((uint8_t)*v);
        } else if (type == 'i' || type == 'I') {
            if (v >= fe) bad("incomplete aux field", line, e);
            if (*v == '-') {
                const int64_t x = parse_int(v, fe);
                if (x >= -128) { o.push_back('c'); o.push_back((uint8_t)(int8_t)x); }
                else if (x >= -32768) { o.push_back('s'); o.push_back((uint8_t)x); o.push_back((uint8_t)(x >> 8)); }
                else { o.push_back('i'); put32(o, (uint32_t)(int32_t)x); }
            } else {
                const uint64_t x = parse_uint(v, fe);
                if (x <= 255) { o.push_back('C'); o.push_back((uint8_t)x); }
                else if (x <= 65535) { o.push_back('S'); o.push_back((uint8_t)x); o.push_back((uint8_t)(x >> 8)); }
                else { o.push_back('I'); put32(o, (uint32_t)x); }
            }
        } else if (type == 'f') {
            std::string tmp(v, fe);
            const float f = (float)strtod(tmp.c_str(), nullptr);
            uint32_t u; memcpy(&u, &f, 4);
            o.push_back('f'); put32(o, u);
        } else if (type == 'Z' || type == 'H') {
            o.push_back((uint8_t)type);
            o.insert(o.end(), v, fe);
            o.push_back(0);
        } else bad("unsupported aux type", p, fe);
        p = t ? t + 1 : e;
    }
    uint8_t *h = o.data() + at;
    set32(h, (uint32_t)(o.size() - at - 4));
    set32(h + 4, (uint32_t)tid);
    set32(h + 8, (uint32_t)(int32_t)pos);
    set32(h + 12, (uint32_t)bin << 16 | mapq << 8 | (uint32_t)(l_qn + 1));
    set32(h + 16, flag << 16 | (n_cigar & 0xffff));
    set32(h + 20, l_seq);
    set32(h + 24, (uint32_t)mtid);
    set32(h + 28, (uint32_t)(int32_t)mpos);
    set32(h + 32, (uint32_t)(int32_t)isize);
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
