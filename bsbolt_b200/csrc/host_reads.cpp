// host_reads.cpp -- see host_reads.h
#include "host_reads.h"
#include <ctype.h>
#include <algorithm>
#include <string.h>
#include <stdio.h>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

#include <stdlib.h>

namespace bsb {

HostAllocHooks g_host_alloc = {malloc, free, nullptr};

enum { SEP_SPACE = 0, SEP_LINE = 2 };
static const int kBuf = 1 << 20;

// Parsing (and gunzip) of each input file runs on its own thread; records reach the batcher in blocks.
struct FastxReader::Prefetch {
    struct Block {
        std::vector<FastxRecord> rec;   // general parser: the records themselves
        std::vector<char> raw;          // fast parser: one piece of the file
        std::vector<RecView> view;      // what the batcher reads, either way
        int n = 0; int status = 0;
    };
    static const int kBlock = 4096;
    static const size_t kPiece = 8u << 20;
    std::mutex m;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Block>> ready, spare;
    std::deque<std::unique_ptr<Block>> held;    // consumed blocks whose records a batch may still point into, oldest first
    uint64_t n_retired = 0, n_released = 0;     // blocks ever moved to `held` / ever recycled from it
    std::unique_ptr<Block> cur;
    int pos = 0;
    bool done = false, stop = false;
    std::thread th;
};

FastxReader::FastxReader(const std::string &path) : buf_(kBuf)
{
    fp_ = path == "-" ? gzdopen(0, "r") : gzopen(path.c_str(), "r");
    if (!fp_) throw std::runtime_error("[E::main_mem] fail to open file `" + path + "'.");
    gzbuffer(fp_, 1 << 20);
    if (path != "-" && gzdirect(fp_) && !getenv("BSB_SLOW_READER")) {   // not compressed: try the fast parser
        raw_ = fopen(path.c_str(), "rb");
        if (raw_ && fseeko(raw_, 0, SEEK_END) != 0) { fclose(raw_); raw_ = nullptr; }   // pipes and the like
        if (raw_) { rewind(raw_); setvbuf(raw_, nullptr, _IONBF, 0); }
    }
    pf_ = new Prefetch;
    pf_->th = std::thread([this] { pump(); });
}

FastxReader::~FastxReader()
{
    if (pf_) {
        { std::lock_guard<std::mutex> l(pf_->m); pf_->stop = true; }
        pf_->cv.notify_all();
        if (pf_->th.joinable()) pf_->th.join();
        delete pf_;
    }
    if (fp_) gzclose(fp_);
    if (raw_) fclose(raw_);
}

void FastxReader::pump()
{
    Prefetch &P = *pf_;
    for (;;) {
        std::unique_ptr<Prefetch::Block> b;
        {
            std::unique_lock<std::mutex> l(P.m);
            P.cv.wait(l, [&] { return P.stop || P.ready.size() < 96; });
            if (P.stop) return;
            if (!P.spare.empty()) { b = std::move(P.spare.front()); P.spare.pop_front(); }
        }
        if (!b) b.reset(new Prefetch::Block);
        b->n = 0; b->status = 0;
        if (!raw_ || !pump_fast_block(b.get())) {
            if (b->rec.size() < (size_t)Prefetch::kBlock) b->rec.resize(Prefetch::kBlock);
            b->view.resize(Prefetch::kBlock);
            while (b->n < Prefetch::kBlock) {
                FastxRecord &k = b->rec[b->n];
                int r = next_raw(k);
                if (r < 0) { b->status = r; break; }
                RecView &v = b->view[b->n];
                v.name = k.name.data(); v.name_l = (uint32_t)k.name.size(); v.cmt = k.comment.data(); v.cmt_l = (uint32_t)k.comment.size();
                v.seq = k.seq.data(); v.seq_l = (uint32_t)k.seq.size(); v.qual = k.qual.data(); v.qual_l = (uint32_t)k.qual.size();
                ++b->n;
            }
        }
        bool last = b->status < 0;
        {
            std::lock_guard<std::mutex> l(P.m);
            P.ready.push_back(std::move(b));
            if (last) P.done = true;
        }
        P.cv.notify_all();
        if (last) return;
    }
}

// Cuts one piece of the file into records. Returns false (block untouched) when the general parser has to take over:
// raw_ is then closed and the zlib stream positioned on the first byte that was not handed out.
bool FastxReader::pump_fast_block(void *block)
{
    Prefetch::Block &b = *static_cast<Prefetch::Block *>(block);
    std::vector<char> &raw = b.raw;
    raw.resize(carry_.size() + Prefetch::kPiece);
    if (!carry_.empty()) memcpy(raw.data(), carry_.data(), carry_.size());
    const size_t got = fread(raw.data() + carry_.size(), 1, Prefetch::kPiece, raw_);
    const size_t len = carry_.size() + got;
    const bool at_eof = got < Prefetch::kPiece;
    carry_.clear();
    b.view.clear();
    const char *base = raw.data(), *end = base + len;
    const char *p = base;
    bool give_up = false;
    while (p < end) {
        const char *l0 = p, *e0, *e1, *e2, *e3;
        if (!(e0 = (const char *)memchr(l0, '\n', end - l0))) break;
        if (!(e1 = (const char *)memchr(e0 + 1, '\n', end - (e0 + 1)))) break;
        if (!(e2 = (const char *)memchr(e1 + 1, '\n', end - (e1 + 1)))) break;
        if (!(e3 = (const char *)memchr(e2 + 1, '\n', end - (e2 + 1)))) break;
        const char *seq = e0 + 1, *qual = e2 + 1;
        const size_t sl = (size_t)(e1 - seq), ql = (size_t)(e3 - qual);
        // anything but "@name[ comment]\nSEQ\n+...\nQUAL\n" with |SEQ| == |QUAL| > 0 and no carriage returns
        if (*l0 != '@' || e0 == l0 + 1 || e1[1] != '+' || sl == 0 || sl != ql || seq[0] == '+' || seq[0] == '>' || seq[0] == '@' ||
            e0[-1] == '\r' || e1[-1] == '\r' || e3[-1] == '\r') { give_up = true; break; }
        const char *nm = l0 + 1, *q = nm;
        while (q < e0 && !isspace((unsigned char)*q)) ++q;
        if (q == nm) { give_up = true; break; }
        RecView v;
        v.name = nm; v.name_l = (uint32_t)(q - nm);
        v.cmt = q < e0 ? q + 1 : e0; v.cmt_l = (uint32_t)(e0 - v.cmt);
        v.seq = seq; v.seq_l = (uint32_t)sl; v.qual = qual; v.qual_l = (uint32_t)ql;
        b.view.push_back(v);
        p = e3 + 1;
    }
    const size_t used = (size_t)(p - base);
    raw_off_ += (int64_t)used;
    if (!give_up && !at_eof) carry_.assign(p, end);          // an incomplete record: finish it with the next piece
    else if (give_up || p < end) {                            // unrecognised record, or a tail without its last newline
        fclose(raw_); raw_ = nullptr;
        gzseek(fp_, raw_off_, SEEK_SET);
        begin_ = end_ = 0; is_eof_ = false; last_char_ = 0;
        if (b.view.empty()) return false;
    }
    b.n = (int)b.view.size();
    if (at_eof && !give_up && p >= end) { b.status = -1; fclose(raw_); raw_ = nullptr; gzseek(fp_, 0, SEEK_END); begin_ = end_ = 0; is_eof_ = true; }
    return true;
}

int FastxReader::next(FastxRecord &r)
{
    Prefetch &P = *pf_;
    for (;;) {
        if (P.cur && P.pos < P.cur->n) {
            const RecView &s = P.cur->view[P.pos++];
            r.name.assign(s.name, s.name_l); r.comment.assign(s.cmt, s.cmt_l); r.seq.assign(s.seq, s.seq_l); r.qual.assign(s.qual, s.qual_l);
            return (int)r.seq.size();
        }
        if (P.cur && P.cur->status < 0) return P.cur->status;
        std::unique_lock<std::mutex> l(P.m);
        if (P.cur) { P.spare.push_back(std::move(P.cur)); P.cv.notify_all(); }
        P.cv.wait(l, [&] { return !P.ready.empty() || P.done; });
        if (P.ready.empty()) return -1;
        P.cur = std::move(P.ready.front());
        P.ready.pop_front();
        P.pos = 0;
        P.cv.notify_all();
    }
}

RecView *FastxReader::next_ptr()
{
    Prefetch &P = *pf_;
    for (;;) {
        if (P.cur && P.pos < P.cur->n) return &P.cur->view[P.pos++];
        if (P.cur && P.cur->status < 0) return nullptr;
        std::unique_lock<std::mutex> l(P.m);
        if (P.cur) { P.held.push_back(std::move(P.cur)); ++P.n_retired; }
        P.cv.wait(l, [&] { return !P.ready.empty() || P.done; });
        if (P.ready.empty()) return nullptr;
        P.cur = std::move(P.ready.front());
        P.ready.pop_front();
        P.pos = 0;
        P.cv.notify_all();
    }
}

uint64_t FastxReader::hold_mark()
{
    std::lock_guard<std::mutex> l(pf_->m);
    return pf_->n_retired;
}

void FastxReader::release_until(uint64_t mark)
{
    Prefetch &P = *pf_;
    std::lock_guard<std::mutex> l(P.m);
    while (P.n_released < mark && !P.held.empty()) {
        P.spare.push_back(std::move(P.held.front()));
        P.held.pop_front();
        ++P.n_released;
    }
    P.cv.notify_all();
}

int FastxReader::getc_()
{
    if (is_eof_ && begin_ >= end_) return -1;
    if (begin_ >= end_) {
        begin_ = 0;
        end_ = gzread(fp_, buf_.data(), kBuf);
        if (end_ <= 0) { end_ = 0; is_eof_ = true; return -1; }
    }
    return (int)buf_[begin_++];
}

int FastxReader::get_until(int delim, std::string &s, int *dret, bool append)
{
    bool gotany = false;
    if (dret) *dret = 0;
    if (!append) s.clear();
    for (;;) {
        int i;
        if (begin_ >= end_) {
            if (!is_eof_) {
                begin_ = 0;
                end_ = gzread(fp_, buf_.data(), kBuf);
                if (end_ <= 0) { end_ = 0; is_eof_ = true; break; }
            } else break;
        }
        if (delim == SEP_LINE) {
            const unsigned char *p = (const unsigned char *)memchr(buf_.data() + begin_, '\n', end_ - begin_);
            i = p ? (int)(p - buf_.data()) : end_;
        } else {
            for (i = begin_; i < end_; ++i) if (isspace(buf_[i])) break;
        }
        gotany = true;
        s.append((const char *)buf_.data() + begin_, i - begin_);
        begin_ = i + 1;
        if (i < end_) { if (dret) *dret = buf_[i]; break; }
    }
    if (!gotany && eof()) return -1;
    if (delim == SEP_LINE && s.size() > 1 && s.back() == '\r') s.pop_back();
    return (int)s.size();
}

int FastxReader::next_raw(FastxRecord &r)
{
    int c;
    if (last_char_ == 0) {
        while ((c = getc_()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        last_char_ = c;
    }
    r.comment.clear(); r.seq.clear(); r.qual.clear();
    if (get_until(SEP_SPACE, r.name, &c, false) < 0) return -1;
    if (c != '\n') get_until(SEP_LINE, r.comment, nullptr, false);
    while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        r.seq.push_back((char)c);
        get_until(SEP_LINE, r.seq, nullptr, true);
    }
    if (c == '>' || c == '@') last_char_ = c;
    if (c != '+') return (int)r.seq.size();
    while ((c = getc_()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while (get_until(SEP_LINE, r.qual, nullptr, true) >= 0 && r.qual.size() < r.seq.size()) {}
    last_char_ = 0;
    if (r.seq.size() != r.qual.size()) return -2;
    return (int)r.seq.size();
}

void ReadBatch::clear()
{
    n = 0; n_bases = 0;
    seq_off.assign(1, 0); name_off.assign(1, 0); cmt_off.assign(1, 0);
    bases.clear(); qual.clear(); has_qual.clear(); names.clear(); comments.clear();
    first.clear(); read_group.clear(); pattern.clear();
}

void ReadBatch::add(const FastxRecord &r, bool keep_comment, int first_, int read_group_, int pattern_)
{
    // l_seq = strlen(seq): an embedded NUL would end the read in the reference as well
    const size_t l = strnlen(r.seq.data(), r.seq.size());
    const size_t o = bases.size();
    if (bases.capacity() < o + l) { bases.reserve((o + l) * 2 + (1 << 20)); qual.reserve((o + l) * 2 + (1 << 20)); }
    bases.resize(o + l); qual.resize(o + l);
    memcpy(bases.data() + o, r.seq.data(), l);
    const bool hq = !r.qual.empty();
    if (hq) memcpy(qual.data() + o, r.qual.data(), l);
    else memset(qual.data() + o, '*', l);
    has_qual.push_back(hq);
    seq_off.push_back((uint32_t)(o + l));
    const size_t no = names.size();
    if (names.capacity() < no + r.name.size()) names.reserve((no + r.name.size()) * 2 + (1 << 16));
    names.resize(no + r.name.size());
    memcpy(names.data() + no, r.name.data(), r.name.size());
    name_off.push_back((uint32_t)names.size());
    if (keep_comment) comments.insert(comments.end(), r.comment.begin(), r.comment.end());
    cmt_off.push_back((uint32_t)comments.size());
    first.push_back((uint8_t)first_); read_group.push_back((uint8_t)read_group_); pattern.push_back((uint8_t)pattern_);
    n_bases += (int64_t)l;
    ++n;
}

void ReadBatch::reserve_like(const ReadBatch &o)
{
    bases.reserve(o.bases.size() + (o.bases.size() >> 3)); qual.reserve(o.qual.size() + (o.qual.size() >> 3));
    names.reserve(o.names.size() + (o.names.size() >> 3));
    size_t k = (size_t)o.n + (o.n >> 3) + 16;
    seq_off.reserve(k); name_off.reserve(k); cmt_off.reserve(k); has_qual.reserve(k); first.reserve(k); read_group.reserve(k); pattern.reserve(k);
}

static int count_base(const char *s, size_t n, char b)
{
    float c = 0;
    size_t l = strnlen(s, n);
    for (size_t i = 0; i < l; ++i) if (s[i] == b) ++c;
    return (int)c;
}

int assess_conversion(const char *s1, size_t l1, const char *s2, size_t l2, int paired_end, float substitution_proportion)
{
    float observed = (float)strnlen(s1, l1);
    float c_count = (float)count_base(s1, l1, 'C');
    float g_count = (float)count_base(s1, l1, 'G');
    if (paired_end) {
        observed += (float)strnlen(s2, l2);
        g_count += (float)count_base(s2, l2, 'C');
        c_count += (float)count_base(s2, l2, 'G');
    }
    float c_prop = c_count / observed;
    float g_prop = g_count / observed;
    float diff = c_prop - g_prop;
    if (diff < 0) diff = (float)(diff * -1.0);
    if (c_prop > substitution_proportion && g_prop > substitution_proportion) return 2;
    else if (c_prop == g_prop) return 2;
    else if (diff < 0.02) return 2;
    else if (c_prop < g_prop) return 0;
    else return 1;
}

static void trim_readno(RecView &v)
{
    size_t l = v.name_l;
    if (l > 2 && v.name[l - 2] == '/' && isdigit((unsigned char)v.name[l - 1])) v.name_l = (uint32_t)(l - 2);
}

void ReadBatch::fill(const std::vector<Entry> &e, bool keep_comment, int n_threads)
{
    clear();
    const size_t m = e.size();
    seq_off.resize(m + 1); name_off.resize(m + 1); cmt_off.resize(m + 1);
    has_qual.resize(m); first.resize(m); read_group.resize(m); pattern.resize(m);
    uint32_t so = 0, no = 0, co = 0;
    for (size_t i = 0; i < m; ++i) {
        seq_off[i] = so; name_off[i] = no; cmt_off[i] = co;
        so += e[i].len; no += e[i].rec->name_l;
        if (keep_comment) co += e[i].rec->cmt_l;
    }
    seq_off[m] = so; name_off[m] = no; cmt_off[m] = co;
    bases.resize(so); qual.resize(so); names.resize(no); comments.resize(co);
    n = (int)m; n_bases = so;
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const RecView &r = *e[i].rec;
            const size_t l = e[i].len;
            memcpy(bases.data() + seq_off[i], r.seq, l);
            const bool hq = r.qual_l != 0;
            if (hq) memcpy(qual.data() + seq_off[i], r.qual, l);
            else memset(qual.data() + seq_off[i], '*', l);
            has_qual[i] = hq;
            memcpy(names.data() + name_off[i], r.name, r.name_l);
            if (keep_comment) memcpy(comments.data() + cmt_off[i], r.cmt, r.cmt_l);
            first[i] = e[i].first; read_group[i] = e[i].read_group; pattern[i] = e[i].pattern;
        }
    };
    if (n_threads <= 1 || m < 4096) { work(0, m); return; }
    std::vector<std::thread> th;
    const size_t chunk = (m + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        size_t lo = t * chunk, hi = lo + chunk < m ? lo + chunk : m;
        if (lo >= hi) break;
        th.emplace_back(work, lo, hi);
    }
    for (auto &x : th) x.join();
}

bool plan_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, int undirectional, float substitution_proportion, BatchPlan &plan)
{
    // serial, no copying: walk the parsers' record blocks, apply the reference's batching rule (bwa.c:73-145)
    std::vector<ReadBatch::Entry> &ents = plan.ents;
    ents.clear();
    int64_t size = 0;
    auto push = [&](RecView *k, int first, int rg, int pattern) {
        uint32_t l = (uint32_t)strnlen(k->seq, k->seq_l);
        ents.push_back(ReadBatch::Entry{k, l, (uint8_t)first, (uint8_t)rg, (uint8_t)pattern});
        size += l;
    };
    RecView *k1, *k2 = nullptr;
    while ((k1 = r1->next_ptr()) != nullptr) {
        if (r2 && (k2 = r2->next_ptr()) == nullptr) {
            fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", "bseq_read");
            break;
        }
        trim_readno(*k1);
        int pattern = 0, compare_reads = 0;
        if (undirectional) {
            int un_type = r2 ? assess_conversion(k1->seq, k1->seq_l, k2->seq, k2->seq_l, 1, substitution_proportion)
                             : assess_conversion(k1->seq, k1->seq_l, k1->seq, k1->seq_l, 0, substitution_proportion);
            if (un_type == 2) compare_reads = 1;
            else pattern = un_type;
        }
        push(k1, 0, 0, pattern);
        if (r2) { trim_readno(*k2); push(k2, 1, 0, pattern ? 0 : 1); }
        if (compare_reads) {
            push(k1, 0, 1, 1);
            if (r2) push(k2, 1, 1, 0);
        }
        if (size >= chunk_size && (ents.size() & 1) == 0) break;
    }
    if (size == 0 && ents.empty()) {
        if (r2 && r1->next_ptr() == nullptr && r2->next_ptr() != nullptr) fprintf(stderr, "[W::%s] the 1st file has fewer sequences.\n", "bseq_read");
    }
    // the block each parser is standing in may also hold records of the next batch: it is not part of this mark
    plan.mark1 = r1->hold_mark();
    plan.mark2 = r2 ? r2->hold_mark() : 0;
    return !ents.empty();
}

void fill_batch(const BatchPlan &plan, FastxReader *r1, FastxReader *r2, bool keep_comment, int n_threads, ReadBatch &b)
{
    b.fill(plan.ents, keep_comment, n_threads);
    r1->release_until(plan.mark1);
    if (r2) r2->release_until(plan.mark2);
}

static int fill_threads()
{
    int nt = (int)std::thread::hardware_concurrency() / 2;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) nt /= std::max(1, atoi(e));
    if (const char *e = getenv("BSB_HOST_THREADS")) nt = atoi(e) / 2;
    return nt < 1 ? 1 : nt > 8 ? 8 : nt;
}

bool read_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, bool keep_comment, int undirectional,
                float substitution_proportion, ReadBatch &b)
{
    static thread_local BatchPlan plan;
    const bool any = plan_batch(chunk_size, r1, r2, undirectional, substitution_proportion, plan);
    fill_batch(plan, r1, r2, keep_comment, fill_threads(), b);
    return any && b.n > 0;
}

int host_fill_threads() { return fill_threads(); }

} // namespace bsb
