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
#include <map>
#include <atomic>

#include <stdlib.h>
#include <unistd.h>
#include <sys/stat.h>
#include <sys/mman.h>

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
        std::vector<uint32_t> cum, cumn, cumc;   // running sums of read / name / comment lengths (view.size() + 1 each)
        int n = 0; int status = 0;
        void begin() { view.clear(); cum.assign(1, 0); cumn.assign(1, 0); cumc.assign(1, 0); n = 0; status = 0; }
        void add(const RecView &v)
        {
            view.push_back(v);
            cum.push_back(cum.back() + v.len); cumn.push_back(cumn.back() + v.name_l); cumc.push_back(cumc.back() + v.cmt_l);
        }
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
    // parallel fast path: pieces [k * piece, (k + 1) * piece) of the file, parsed out of order, chained in order
    struct Piece {
        std::unique_ptr<Block> blk;
        int64_t first_off = -1, end_off = -1;   // file offsets of the first record's '@' / just past the last record
        bool give_up = false;                   // the record at end_off is not strict four-line FASTQ (or has no final newline)
        bool unsynced = false;                  // no record boundary could be recognised in this piece
    };
    size_t piece = kPiece;
    std::map<int64_t, Piece> parsed;            // finished pieces waiting for their turn
    int64_t next_claim = 0, next_deliver = 0, n_pieces = 0;
    bool par_stop = false;
    std::unique_ptr<Block> take_spare()
    {
        std::unique_ptr<Block> b;
        std::lock_guard<std::mutex> l(m);
        if (!spare.empty()) { b = std::move(spare.front()); spare.pop_front(); }
        return b;
    }
};

// what the batcher needs of every record, computed where the parsing is parallel: the read as the reference sees it
// (strlen), its C and G counts (conversion-pattern assessment of undirectional libraries) and the name without "/1" "/2"
// bytes of `w` equal to the byte replicated in `pat`, counted eight at a time (exact zero-byte test, no carries between bytes)
static inline uint32_t count_eq8(uint64_t w, uint64_t pat)
{
    const uint64_t x = w ^ pat, m = 0x7f7f7f7f7f7f7f7full;
    return (uint32_t)__builtin_popcountll(~(((x & m) + m) | x | m));
}

static inline void finish_view(RecView &v, bool count_cg)
{
    const uint32_t l = (uint32_t)strnlen(v.seq, v.seq_l);
    uint32_t c = 0, g = 0;
    if (count_cg) {   // only undirectional libraries look at the base composition (assessConversion)
        uint32_t i = 0;
        for (; i + 8 <= l; i += 8) {
            uint64_t w;
            memcpy(&w, v.seq + i, 8);
            c += count_eq8(w, 0x4343434343434343ull); g += count_eq8(w, 0x4747474747474747ull);
        }
        for (; i < l; ++i) { c += v.seq[i] == 'C'; g += v.seq[i] == 'G'; }
    }
    v.len = l; v.n_c = c; v.n_g = g;
    const size_t n = v.name_l;
    if (n > 2 && v.name[n - 2] == '/' && isdigit((unsigned char)v.name[n - 1])) v.name_l = (uint32_t)(n - 2);
}

// Cuts strict four-line FASTQ records out of [p, end) while the record's header starts before `limit`. Stops at an
// incomplete record (returns with *p on its header). give_up: the record at *p is something else -- anything but
// "@name[ comment]\nSEQ\n+...\nQUAL\n" with |SEQ| == |QUAL| > 0 and no carriage returns.
static inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }   // isspace() in the C locale

template <class Blk>
static void cut_records(const char *&p, const char *end, const char *limit, Blk &blk, bool &give_up, bool count_cg)
{
    while (p < limit) {
        const char *l0 = p, *e0, *e1, *e2, *e3;
        if (!(e0 = (const char *)memchr(l0, '\n', end - l0))) break;
        if (!(e1 = (const char *)memchr(e0 + 1, '\n', end - (e0 + 1)))) break;
        if (!(e2 = (const char *)memchr(e1 + 1, '\n', end - (e1 + 1)))) break;
        if (!(e3 = (const char *)memchr(e2 + 1, '\n', end - (e2 + 1)))) break;
        const char *seq = e0 + 1, *qual = e2 + 1;
        const size_t sl = (size_t)(e1 - seq), ql = (size_t)(e3 - qual);
        if (*l0 != '@' || e0 == l0 + 1 || e1[1] != '+' || sl == 0 || sl != ql || seq[0] == '+' || seq[0] == '>' || seq[0] == '@' ||
            e0[-1] == '\r' || e1[-1] == '\r' || e3[-1] == '\r') { give_up = true; break; }
        const char *nm = l0 + 1, *q = nm;
        while (q < e0 && !is_space((unsigned char)*q)) ++q;
        if (q == nm) { give_up = true; break; }
        RecView v;
        v.name = nm; v.name_l = (uint32_t)(q - nm);
        v.cmt = q < e0 ? q + 1 : e0; v.cmt_l = (uint32_t)(e0 - v.cmt);
        v.seq = seq; v.seq_l = (uint32_t)sl; v.qual = qual; v.qual_l = (uint32_t)ql;
        finish_view(v, count_cg);
        blk.add(v);
        p = e3 + 1;
    }
}

FastxReader::FastxReader(const std::string &path, int n_threads, bool count_cg) : buf_(kBuf), n_threads_(n_threads), count_cg_(count_cg)
{
    fp_ = path == "-" ? gzdopen(0, "r") : gzopen(path.c_str(), "r");
    if (!fp_) throw std::runtime_error("[E::main_mem] fail to open file `" + path + "'.");
    gzbuffer(fp_, 1 << 20);
    if (path != "-" && gzdirect(fp_) && !getenv("BSB_SLOW_READER")) {   // not compressed: try the fast parser
        raw_ = fopen(path.c_str(), "rb");
        if (raw_ && fseeko(raw_, 0, SEEK_END) != 0) { fclose(raw_); raw_ = nullptr; }   // pipes and the like
        if (raw_) { file_size_ = (int64_t)ftello(raw_); rewind(raw_); setvbuf(raw_, nullptr, _IONBF, 0); }
    }
    pf_ = new Prefetch;
    if (const char *e = getenv("BSB_READ_PIECE")) pf_->piece = (size_t)std::max(64, atoi(e));   // tests: many pieces in a small file
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
    if (map_) munmap(const_cast<char *>(map_), (size_t)file_size_);   // after the parser threads: records point into the mapping
}

void FastxReader::deliver(void *block)
{
    Prefetch &P = *pf_;
    std::unique_ptr<Prefetch::Block> b(static_cast<Prefetch::Block *>(block));
    const bool last = b->status < 0;
    {
        std::lock_guard<std::mutex> l(P.m);
        P.ready.push_back(std::move(b));
        if (last) P.done = true;
    }
    P.cv.notify_all();
}

// One piece of the file, on any thread: finds the first record header at or after the piece's first byte by its
// signature (a line starting with '@' whose next-but-one line starts with '+': in strict four-line FASTQ, where no
// sequence starts with '@' or '+', only a header line matches), then cuts records until one starts in the next piece.
void FastxReader::parse_piece(void *piece, int64_t k)
{
    Prefetch &P = *pf_;
    Prefetch::Piece &pc = *static_cast<Prefetch::Piece *>(piece);
    pc.blk = P.take_spare();
    if (!pc.blk) pc.blk.reset(new Prefetch::Block);
    Prefetch::Block &b = *pc.blk;
    b.begin();
    const int64_t lo = k * (int64_t)P.piece, hi = std::min<int64_t>(lo + (int64_t)P.piece, file_size_);
    // the file is mapped: records are cut where they lie in the page cache (a record that starts in this piece may run on into
    // the next), nothing is copied until the batch is filled -- at eight GPUs the host's memory bandwidth is what the reader
    // runs out of, and every copy of the input that is not made counts
    const char *base = map_, *end = map_ + file_size_, *limit = map_ + hi;
    const char *p = map_ + lo;
    if (lo > 0) {
        if (p[-1] != '\n') { const char *nl = (const char *)memchr(p, '\n', end - p); p = nl ? nl + 1 : end; }   // first line start at or after lo
        bool found = false, short_of_data = false;
        for (int tries = 0; tries < 6 && p < limit; ++tries) {
            const char *e0 = (const char *)memchr(p, '\n', end - p);
            const char *e1 = e0 ? (const char *)memchr(e0 + 1, '\n', end - (e0 + 1)) : nullptr;
            if (!e1 || e1 + 1 >= end) { short_of_data = true; break; }
            if (*p == '@' && e1[1] == '+') { found = true; break; }
            p = e0 + 1;
        }
        if (!found && p < limit && !short_of_data) { pc.unsynced = true; return; }
        // no complete record starts in this piece (or, at the end of the file, only a truncated one: the chain check of the
        // consumer then sends the serial parser there)
        if (!found) { pc.first_off = pc.end_off = -1; return; }
    }
    const char *first = p;
    bool give_up = false;
    cut_records(p, end, limit, b, give_up, count_cg_);
    pc.first_off = first - base;
    pc.end_off = p - base;
    pc.give_up = give_up || (p < limit);           // at the end of the file: a tail without its final newline
    b.n = (int)b.view.size();
}

bool FastxReader::pump_parallel()
{
    Prefetch &P = *pf_;
    void *mp = mmap(nullptr, (size_t)file_size_, PROT_READ, MAP_PRIVATE | MAP_NORESERVE, fileno(raw_), 0);
    if (mp == MAP_FAILED) return false;            // the serial parser reads the file from its start
    map_ = static_cast<const char *>(mp);
    madvise(mp, (size_t)file_size_, MADV_SEQUENTIAL);

    P.n_pieces = (file_size_ + (int64_t)P.piece - 1) / (int64_t)P.piece;
    const int window = 2 * n_threads_ + 2;                     // pieces parsed ahead of the one being handed out
    std::vector<std::thread> workers;
    for (int t = 0; t < n_threads_; ++t)
        workers.emplace_back([&] {
            for (;;) {
                int64_t k;
                {
                    std::unique_lock<std::mutex> l(P.m);
                    P.cv.wait(l, [&] { return P.stop || P.par_stop || P.next_claim >= P.n_pieces || P.next_claim < P.next_deliver + window; });
                    if (P.stop || P.par_stop || P.next_claim >= P.n_pieces) return;
                    k = P.next_claim++;
                }
                Prefetch::Piece pc;
                parse_piece(&pc, k);
                {
                    std::lock_guard<std::mutex> l(P.m);
                    P.parsed[k] = std::move(pc);
                }
                P.cv.notify_all();
            }
        });
    int64_t expect = 0;                                        // file offset the next record must start at
    bool whole = true, sent_last = false;
    for (int64_t k = 0; k < P.n_pieces; ++k) {
        Prefetch::Piece pc;
        {
            std::unique_lock<std::mutex> l(P.m);
            P.cv.wait(l, [&] { return P.stop || P.parsed.count(k); });
            if (P.stop) { whole = true; break; }
            pc = std::move(P.parsed[k]);
            P.parsed.erase(k);
            P.cv.wait(l, [&] { return P.stop || P.ready.size() < 24; });
            if (P.stop) break;
            P.next_deliver = k + 1;
        }
        P.cv.notify_all();
        const int64_t lo = k * (int64_t)P.piece, hi = std::min<int64_t>(lo + (int64_t)P.piece, file_size_);
        if (pc.first_off < 0 && !pc.unsynced) {                // nothing starts in this piece: the previous record must cover it
            if (expect < hi) { whole = false; break; }
            { std::lock_guard<std::mutex> l(P.m); P.spare.push_back(std::move(pc.blk)); }
            continue;
        }
        if (pc.unsynced || pc.first_off != expect) { whole = false; break; }
        expect = pc.end_off;
        const bool last = k + 1 == P.n_pieces;
        if (pc.give_up) {                                      // hand out what was cut, the general parser continues at `expect`
            if (pc.blk->n) deliver(pc.blk.release());
            whole = false;
            break;
        }
        if (last) { pc.blk->status = -1; sent_last = true; }
        if (pc.blk->n || last) deliver(pc.blk.release());
        else { std::lock_guard<std::mutex> l(P.m); P.spare.push_back(std::move(pc.blk)); }
    }
    {
        std::lock_guard<std::mutex> l(P.m);
        P.par_stop = true;
    }
    P.cv.notify_all();
    for (auto &w : workers) w.join();
    {
        std::lock_guard<std::mutex> l(P.m);
        for (auto &kv : P.parsed) if (kv.second.blk) P.spare.push_back(std::move(kv.second.blk));
        P.parsed.clear();
    }
    if (getenv("BSB_DEBUG_READER"))
        fprintf(stderr, "[D::reader] %d threads, %ld pieces of %zu bytes: %s at offset %ld of %ld\n", n_threads_, (long)P.n_pieces, P.piece,
                whole ? "whole file cut in parallel" : "serial parser takes over", (long)expect, (long)file_size_);
    if (whole) {
        if (!sent_last) { std::unique_ptr<Prefetch::Block> b(new Prefetch::Block); b->status = -1; deliver(b.release()); }
        fclose(raw_); raw_ = nullptr; gzseek(fp_, 0, SEEK_END); begin_ = end_ = 0; is_eof_ = true;
        return true;
    }
    raw_off_ = expect;
    fseeko(raw_, (off_t)expect, SEEK_SET);
    return false;
}

void FastxReader::pump()
{
    Prefetch &P = *pf_;
    if (raw_ && file_size_ > 0 && pump_parallel()) return;
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
            b->begin();
            while (b->n < Prefetch::kBlock) {
                FastxRecord &k = b->rec[b->n];
                int r = next_raw(k);
                if (r < 0) { b->status = r; break; }
                RecView v;
                v.name = k.name.data(); v.name_l = (uint32_t)k.name.size(); v.cmt = k.comment.data(); v.cmt_l = (uint32_t)k.comment.size();
                v.seq = k.seq.data(); v.seq_l = (uint32_t)k.seq.size(); v.qual = k.qual.data(); v.qual_l = (uint32_t)k.qual.size();
                finish_view(v, count_cg_);
                b->add(v);
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
    b.begin();
    const char *base = raw.data(), *end = base + len;
    const char *p = base;
    bool give_up = false;
    cut_records(p, end, end, b, give_up, count_cg_);
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

bool FastxReader::run(RecRun &r)
{
    Prefetch &P = *pf_;
    for (;;) {
        if (P.cur && P.pos < P.cur->n) {
            const Prefetch::Block &b = *P.cur;
            r.v = b.view.data() + P.pos; r.cum = b.cum.data() + P.pos; r.cumn = b.cumn.data() + P.pos; r.cumc = b.cumc.data() + P.pos;
            r.n = b.n - P.pos;
            return true;
        }
        if (P.cur && P.cur->status < 0) return false;
        std::unique_lock<std::mutex> l(P.m);
        if (P.cur) { P.held.push_back(std::move(P.cur)); ++P.n_retired; }
        P.cv.wait(l, [&] { return !P.ready.empty() || P.done; });
        if (P.ready.empty()) return false;
        P.cur = std::move(P.ready.front());
        P.ready.pop_front();
        P.pos = 0;
        P.cv.notify_all();
    }
}

void FastxReader::consume(int k) { pf_->pos += k; }

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

// assessConversion (bs_helpers.cpp:41-62) in the reference's float arithmetic, from the parser's counts
int assess_conversion_counts(const RecView &k1, const RecView *k2, float substitution_proportion)
{
    float observed = (float)k1.len, c_count = (float)(int)(float)k1.n_c, g_count = (float)(int)(float)k1.n_g;
    if (k2) {
        observed += (float)k2->len;
        g_count += (float)(int)(float)k2->n_c;
        c_count += (float)(int)(float)k2->n_g;
    }
    const float c_prop = c_count / observed, g_prop = g_count / observed;
    float diff = c_prop - g_prop;
    if (diff < 0) diff = (float)(diff * -1.0);
    if (c_prop > substitution_proportion && g_prop > substitution_proportion) return 2;
    else if (c_prop == g_prop) return 2;
    else if (diff < 0.02) return 2;
    else if (c_prop < g_prop) return 0;
    else return 1;
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

// Directional libraries: every record is exactly one entry with conversion pattern 0 (first mates / single end) or 1
// (second mates), so the batch is cut through the blocks' running sums -- whole runs of records per step, a binary search in
// the run where the reference's rule (bwa.c:73-145: stop after the pair that brings the batch to chunk_size bases, on an even
// number of entries) fires -- and the entries themselves are only spelled out by fill_segments(), in parallel.
static bool plan_batch_segments(int64_t chunk_size, FastxReader *r1, FastxReader *r2, BatchPlan &plan)
{
    plan.segs.clear(); plan.ents.clear();
    plan.by_segments = true; plan.paired = r2 != nullptr;
    int64_t size = 0, n_ent = 0;
    RecRun a, b;
    for (;;) {
        if (!r1->run(a)) break;
        if (r2 && !r2->run(b)) {
            fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", "bseq_read");
            r1->consume(1);                        // the reference has read this record and drops it
            break;
        }
        const int run = r2 ? (a.n < b.n ? a.n : b.n) : a.n;
        auto bases = [&](int k) { return (int64_t)(a.cum[k] - a.cum[0]) + (r2 ? (int64_t)(b.cum[k] - b.cum[0]) : 0); };
        int take = run;
        if (size + bases(run) >= chunk_size) {     // the rule fires inside this run: first k with size + bases(k) >= chunk_size
            int lo = 1, hi = run;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (size + bases(mid) >= chunk_size) hi = mid; else lo = mid + 1; }
            take = lo;
            if (!r2 && ((n_ent + take) & 1) && take < run) ++take;   // single end: the entry count must be even as well
        }
        ReadBatch::Seg sg; sg.a = a; sg.b = b; sg.n = take;
        plan.segs.push_back(sg);
        size += bases(take); n_ent += r2 ? 2 * (int64_t)take : take;
        r1->consume(take);
        if (r2) r2->consume(take);
        if (size >= chunk_size && (n_ent & 1) == 0) break;
    }
    if (size == 0 && plan.segs.empty()) {
        RecRun t;
        if (r2 && !r1->run(t) && r2->run(t)) fprintf(stderr, "[W::%s] the 1st file has fewer sequences.\n", "bseq_read");
    }
    plan.mark1 = r1->hold_mark();
    plan.mark2 = r2 ? r2->hold_mark() : 0;
    plan.n_entries = n_ent;
    return !plan.segs.empty();
}

void ReadBatch::fill_segments(const std::vector<Seg> &segs, bool paired, bool keep_comment, int n_threads)
{
    clear();
    // where each segment starts in the flat arrays (serial over the handful of segments)
    struct Start { size_t rec, seq, name, cmt; };
    std::vector<Start> st(segs.size() + 1);
    Start cur = {0, 0, 0, 0};
    for (size_t k = 0; k < segs.size(); ++k) {
        st[k] = cur;
        const Seg &g = segs[k];
        cur.rec += (size_t)g.n;
        cur.seq += g.a.cum[g.n] - g.a.cum[0]; cur.name += g.a.cumn[g.n] - g.a.cumn[0];
        if (keep_comment) cur.cmt += g.a.cumc[g.n] - g.a.cumc[0];
        if (paired) {
            cur.seq += g.b.cum[g.n] - g.b.cum[0]; cur.name += g.b.cumn[g.n] - g.b.cumn[0];
            if (keep_comment) cur.cmt += g.b.cumc[g.n] - g.b.cumc[0];
        }
    }
    st[segs.size()] = cur;
    const size_t n_rec = cur.rec, m = paired ? 2 * n_rec : n_rec;
    if (cur.seq > 0xfffffff0ull || cur.name > 0xfffffff0ull) throw std::runtime_error("[E::bseq_read] a batch of 4 GiB of bases or more: lower -K");
    seq_off.resize(m + 1); name_off.resize(m + 1); cmt_off.resize(m + 1);
    has_qual.resize(m); first.resize(m); read_group.resize(m); pattern.resize(m);
    seq_off[m] = (uint32_t)cur.seq; name_off[m] = (uint32_t)cur.name; cmt_off[m] = (uint32_t)cur.cmt;
    bases.resize(cur.seq); qual.resize(cur.seq); names.resize(cur.name); comments.resize(cur.cmt);
    n = (int)m; n_bases = (int64_t)cur.seq;
    auto put = [&](size_t e, const RecView &r, uint32_t so, uint32_t no, uint32_t co, int fst) {
        seq_off[e] = so; name_off[e] = no; cmt_off[e] = co;
        memcpy(bases.data() + so, r.seq, r.len);
        const bool hq = r.qual_l != 0;
        if (hq) memcpy(qual.data() + so, r.qual, r.len);
        else memset(qual.data() + so, '*', r.len);
        has_qual[e] = hq;
        memcpy(names.data() + no, r.name, r.name_l);
        if (keep_comment) memcpy(comments.data() + co, r.cmt, r.cmt_l);
        first[e] = (uint8_t)fst; read_group[e] = 0; pattern[e] = (uint8_t)fst;   // first mates C->T (0), second mates G->A (1)
    };
    auto work = [&](size_t lo, size_t hi) {                // records [lo, hi) of the batch
        size_t k = 0;
        while (k + 1 < segs.size() && st[k + 1].rec <= lo) ++k;
        for (size_t rec = lo; rec < hi;) {
            const Seg &g = segs[k];
            const size_t p0 = rec - st[k].rec, p1 = std::min<size_t>((size_t)g.n, hi - st[k].rec);
            for (size_t p = p0; p < p1; ++p) {
                uint32_t so = (uint32_t)(st[k].seq + (g.a.cum[p] - g.a.cum[0])), no = (uint32_t)(st[k].name + (g.a.cumn[p] - g.a.cumn[0]));
                uint32_t co = (uint32_t)(st[k].cmt + (keep_comment ? g.a.cumc[p] - g.a.cumc[0] : 0));
                if (paired) {
                    so += g.b.cum[p] - g.b.cum[0]; no += g.b.cumn[p] - g.b.cumn[0];
                    if (keep_comment) co += g.b.cumc[p] - g.b.cumc[0];
                    const size_t e = 2 * (st[k].rec + p);
                    put(e, g.a.v[p], so, no, co, 0);
                    put(e + 1, g.b.v[p], so + g.a.v[p].len, no + g.a.v[p].name_l, co + (keep_comment ? g.a.v[p].cmt_l : 0), 1);
                } else put(st[k].rec + p, g.a.v[p], so, no, co, 0);
            }
            rec = st[k].rec + p1;
            ++k;
        }
    };
    if (n_threads <= 1 || n_rec < 4096) { work(0, n_rec); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n_rec + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const size_t lo = t * chunk, hi = lo + chunk < n_rec ? lo + chunk : n_rec;
        if (lo >= hi) break;
        th.emplace_back(work, lo, hi);
    }
    for (auto &x : th) x.join();
}

bool plan_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, int undirectional, float substitution_proportion, BatchPlan &plan)
{
    if (!undirectional && !getenv("BSB_PLAN_ENTRIES")) return plan_batch_segments(chunk_size, r1, r2, plan);
    plan.by_segments = false; plan.segs.clear();
    // serial, no copying: walk the parsers' record blocks, apply the reference's batching rule (bwa.c:73-145)
    std::vector<ReadBatch::Entry> &ents = plan.ents;
    ents.clear();
    int64_t size = 0;
    auto push = [&](RecView *k, int first, int rg, int pattern) {
        ents.push_back(ReadBatch::Entry{k, k->len, (uint8_t)first, (uint8_t)rg, (uint8_t)pattern});
        size += k->len;
    };
    RecView *k1, *k2 = nullptr;
    while ((k1 = r1->next_ptr()) != nullptr) {
        if (r2 && (k2 = r2->next_ptr()) == nullptr) {
            fprintf(stderr, "[W::%s] the 2nd file has fewer sequences.\n", "bseq_read");
            break;
        }
        int pattern = 0, compare_reads = 0;
        if (undirectional) {
            const int un_type = assess_conversion_counts(*k1, r2 ? k2 : nullptr, substitution_proportion);
            if (un_type == 2) compare_reads = 1;
            else pattern = un_type;
        }
        push(k1, 0, 0, pattern);
        if (r2) push(k2, 1, 0, pattern ? 0 : 1);
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
    plan.n_entries = (int64_t)ents.size();
    return !ents.empty();
}

void fill_batch(const BatchPlan &plan, FastxReader *r1, FastxReader *r2, bool keep_comment, int n_threads, ReadBatch &b)
{
    if (plan.by_segments) b.fill_segments(plan.segs, plan.paired, keep_comment, n_threads);
    else b.fill(plan.ents, keep_comment, n_threads);
    r1->release_until(plan.mark1);
    if (r2) r2->release_until(plan.mark2);
}

int host_core_share()
{
    int n = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("BSB_HOST_THREADS")) n = atoi(e);
    else if (const char *e = getenv("LOCAL_WORLD_SIZE")) { if (!getenv("BSB_ALL_CORES")) n /= std::max(1, atoi(e)); }
    return n < 1 ? 1 : n;
}

static int fill_threads(int n_devices = 1)
{
    if (const char *e = getenv("BSB_FILL_THREADS")) return std::max(1, atoi(e));
    const int nt = host_core_share() / 2, cap = n_devices > 2 ? 16 : 8;
    return nt < 1 ? 1 : nt > cap ? cap : nt;
}

// parser threads per input file: one keeps a GPU fed (about 18 M records/s); more devices, more pieces in flight
int host_parse_threads(int n_devices)
{
    if (const char *e = getenv("BSB_PARSE_THREADS")) return std::max(1, atoi(e));
    const int share = host_core_share() / 4;
    const int want = n_devices + 1 < 5 ? n_devices + 1 : 5;   // a thread cuts about 8 M records/s out of the mapped file; a GPU takes 6 M per file
    return std::max(1, std::min(want, std::max(share, 1)));
}

bool read_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, bool keep_comment, int undirectional,
                float substitution_proportion, ReadBatch &b)
{
    static thread_local BatchPlan plan;
    const bool any = plan_batch(chunk_size, r1, r2, undirectional, substitution_proportion, plan);
    fill_batch(plan, r1, r2, keep_comment, fill_threads(), b);
    return any && b.n > 0;
}

int host_fill_threads(int n_devices) { return fill_threads(n_devices); }

} // namespace bsb
