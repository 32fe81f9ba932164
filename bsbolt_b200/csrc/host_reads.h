// host_reads.h -- FASTQ/FASTA (optionally gzip) reader and the batcher that cuts the input into the
// same batches as the reference, including the bisulfite conversion-pattern bookkeeping.
//   FastxReader::next()  <- kseq_read              (kseq.h:175-217)
//   assess_conversion()  <- assessConversion       (bs_helpers.cpp:41-62)
//   read_batch()         <- bseq_read, kseq2bseq1  (bwa.c:44-145)
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>

namespace bsb {

// Host buffers that cross PCIe are allocated through these hooks: malloc/free by default, page-locked memory
// once the CUDA library installs its allocator (bsb_cuda.cu). Set once, before any buffer exists.
struct HostAllocHooks {
    void *(*alloc)(size_t); void (*release)(void *);
    // optional: make sure that, for every i, count[i] blocks able to hold bytes[i] exist (in use or on hand); requests
    // that fall into the same size class add up
    // free_only: count[i] blocks must be on hand (not counting those handed out)
    void (*prefill)(const size_t *bytes, const int *count, int n, bool free_only);
};
extern HostAllocHooks g_host_alloc;

// growable byte buffer without value-initialisation (a 100 MB std::vector::resize costs ~10 ms of memset)
class RawBuf {
public:
    explicit RawBuf(bool transfer_buffer = true) : hooks_(transfer_buffer) {}
    ~RawBuf() { if (p_) rel(p_); }
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    uint8_t *data() { return p_; }
    const uint8_t *data() const { return p_; }
    size_t size() const { return n_; }
    void reset() { if (p_) rel(p_); p_ = nullptr; n_ = cap_ = 0; }   // hand the block back (to the page-locked pool)
    void resize_uninit(size_t n)
    {
        if (n > cap_) {
            size_t c = n + n / 4 + 4096;
            if (p_) rel(p_);                      // contents are not preserved: callers refill the buffer
            p_ = (uint8_t *)(hooks_ ? g_host_alloc.alloc(c) : malloc(c));
            cap_ = c;
        }
        n_ = n;
    }
private:
    void rel(void *p) { if (hooks_) g_host_alloc.release(p); else free(p); }
    uint8_t *p_ = nullptr; size_t n_ = 0, cap_ = 0;
    bool hooks_;
};

// fixed-type array over a RawBuf (contents undefined after resize)
template <class T>
class PinArray {
public:
    T *data() { return reinterpret_cast<T *>(buf_.data()); }
    const T *data() const { return reinterpret_cast<const T *>(buf_.data()); }
    T &operator[](size_t i) { return data()[i]; }
    const T &operator[](size_t i) const { return data()[i]; }
    size_t size() const { return n_; }
    void resize(size_t n) { buf_.resize_uninit(n * sizeof(T)); n_ = n; }
    void assign(size_t n, const T &v) { resize(n); for (size_t i = 0; i < n; ++i) data()[i] = v; }
    void reset() { buf_.reset(); n_ = 0; }
private:
    RawBuf buf_; size_t n_ = 0;
};

// std::vector<char>-like, hook-allocated (page-locked in the product), contents preserved on growth
class PinVec {
public:
    PinVec() {}
    ~PinVec() { if (p_) g_host_alloc.release(p_); }
    PinVec(const PinVec &) = delete;
    PinVec &operator=(const PinVec &) = delete;
    char *data() { return p_; }
    const char *data() const { return p_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    void clear() { n_ = 0; }
    void reset() { if (p_) g_host_alloc.release(p_); p_ = nullptr; n_ = cap_ = 0; }
    void reserve(size_t c)
    {
        if (c <= cap_) return;
        char *q = (char *)g_host_alloc.alloc(c);
        if (p_) { memcpy(q, p_, n_); g_host_alloc.release(p_); }
        p_ = q; cap_ = c;
    }
    void resize(size_t n) { if (n > cap_) reserve(n + n / 2 + 4096); n_ = n; }
private:
    char *p_ = nullptr; size_t n_ = 0, cap_ = 0;
};

struct FastxRecord { std::string name, comment, seq, qual; };

// One parsed record as the batcher sees it: spans into the parser's block (text of the file itself on the fast path)
struct RecView {
    const char *name, *cmt, *seq, *qual;
    uint32_t name_l, cmt_l, seq_l, qual_l;   // qual_l == 0: no quality string; name_l: "/1" "/2" already trimmed (bwa.c:28-32)
    uint32_t len;                            // strlen(seq): what the reference takes as the read (bwa.c:44-57)
    uint32_t n_c, n_g;                       // 'C's and 'G's among those `len` bases (countBase, bs_helpers.cpp:31-39)
};

// A run of parsed records inside one parser block, with running sums over the records (cum*[k] = sum over the first k records
// of the run's block prefix; only differences are meaningful): read length, name length, comment length. The batcher cuts and
// addresses a batch through these sums without touching the records one by one.
struct RecRun {
    const RecView *v = nullptr;
    const uint32_t *cum = nullptr, *cumn = nullptr, *cumc = nullptr;
    int n = 0;
};

// Two parsers behind one interface, each on its own thread, handing blocks of records to the batcher:
//  * fast: plain (uncompressed, seekable) files made of strict four-line FASTQ records are read in multi-megabyte
//    pieces and cut with memchr; records are spans into the piece, nothing is copied until the batch is filled;
//  * general: kseq semantics (kseq.h:175-217) over zlib -- gzip, stdin, FASTA, multi-line records. The fast parser
//    hands over to it at the byte offset of the first record it does not recognise, so results never differ.
class FastxReader {
public:
    // n_threads > 1: a plain four-line FASTQ file is cut by that many threads at once (pieces of the file parsed
    // speculatively from a record boundary found by its "@...\n...\n+" signature, then chained in file order: a piece
    // is only handed out when it starts exactly where its predecessor ended, else the serial parser takes over there)
    // count_cg: fill RecView::n_c / n_g (only undirectional libraries need them)
    explicit FastxReader(const std::string &path, int n_threads = 1, bool count_cg = true);
    ~FastxReader();
    // zero-copy form for the batcher: the next record inside the parser's block (nullptr at end of input). The
    // pointer stays valid until release_held().
    RecView *next_ptr();
    // block-wise form: the records of the current block that have not been consumed yet (false at the end of the input);
    // consume(k) takes the first k of them. Pointers stay valid until release_held() like those of next_ptr().
    bool run(RecRun &r);
    void consume(int k);
    // Blocks the batcher has walked past stay alive until released. hold_mark() names the blocks fully consumed so
    // far; release_until(mark) recycles them (callable from another thread than next_ptr()'s).
    uint64_t hold_mark();
    void release_until(uint64_t mark);
    void release_held() { release_until(hold_mark()); }
private:
    int next_raw(FastxRecord &r);
    void pump();               // parser thread body
    bool pump_fast_block(void *block);
    bool pump_parallel();      // true: the whole file was handed out; false: continue serially at raw_off_
    void parse_piece(void *piece, int64_t k);
    void deliver(void *block);
    struct Prefetch;           // blocks of parsed records handed over from the parser thread
    Prefetch *pf_ = nullptr;
    int getc_();
    int get_until(int delim, std::string &s, int *dret, bool append);
    bool eof() const { return is_eof_ && begin_ >= end_; }
    gzFile fp_;
    std::vector<unsigned char> buf_;
    int begin_ = 0, end_ = 0;
    bool is_eof_ = false;
    int last_char_ = 0;
    // fast path
    FILE *raw_ = nullptr;      // non-null while the fast parser is active
    int64_t raw_off_ = 0;      // file offset of the first byte not yet handed out as a record
    std::vector<char> carry_;  // incomplete record at the end of the previous piece
    int n_threads_ = 1;
    bool count_cg_ = true;
    int64_t file_size_ = 0;
    const char *map_ = nullptr;   // the plain-FASTQ file, mapped (multi-threaded cutter)
};

// One batch of bseq entries (a read aligned under both conversion patterns appears twice).
struct ReadBatch {
    int n = 0;
    std::vector<uint32_t> seq_off;   // n+1 offsets into bases/qual
    PinVec bases;                    // ASCII as read from the file (unconverted); uploaded to the device
    PinVec qual;                     // same offsets as bases; valid iff has_qual (uploaded for the device SAM formatter)
    std::vector<uint8_t> has_qual;
    std::vector<uint32_t> name_off;  // n+1
    PinVec names;
    std::vector<uint32_t> cmt_off;   // n+1 (empty ranges unless -C)
    std::vector<char> comments;
    std::vector<uint8_t> first, read_group, pattern;
    int64_t n_bases = 0;
    void *dev_input = nullptr;       // set by BatchAligner::preload(): the batch's inputs are already resident on the device
    void *dev_owner = nullptr;       // ... of this aligner (multi-GPU runs)

    void clear();
    void reserve_like(const ReadBatch &o);   // pre-size for a batch about as large as o
    struct Entry { const RecView *rec; uint32_t len; uint8_t first, read_group, pattern; };
    void fill(const std::vector<Entry> &e, bool keep_comment, int n_threads);   // bulk, multi-threaded add()
    // n consecutive records of file 1 (a) and, when paired, their mates in file 2 (b): one directional-library entry each
    struct Seg { RecRun a, b; int n; };
    void fill_segments(const std::vector<Seg> &segs, bool paired, bool keep_comment, int n_threads);
    void add(const FastxRecord &r, bool keep_comment, int first, int read_group, int pattern);
    int len(int i) const { return (int)(seq_off[i + 1] - seq_off[i]); }
    std::string name(int i) const { return std::string(names.data() + name_off[i], name_off[i + 1] - name_off[i]); }
};

int assess_conversion(const char *s1, size_t l1, const char *s2, size_t l2, int paired_end, float substitution_proportion);
// the same decision from the parser's per-record counts (RecView::len, n_c, n_g)
int assess_conversion_counts(const RecView &k1, const RecView *k2, float substitution_proportion);

// The two halves of read_batch(), so that cutting batch i+1 can overlap with copying batch i: plan_batch() applies the
// reference's batching rule and records which parser records make up the batch; fill_batch() copies them into the flat
// arrays (multi-threaded) and lets the parsers recycle the blocks.
struct BatchPlan {
    std::vector<ReadBatch::Entry> ents;    // undirectional libraries: the entries one by one (a read may appear twice)
    std::vector<ReadBatch::Seg> segs;      // directional libraries: runs of records, cut and addressed through their running sums
    bool by_segments = false, paired = false;
    int64_t n_entries = 0;                 // bseq entries of the batch, either way
    uint64_t mark1 = 0, mark2 = 0;
};
bool plan_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, int undirectional, float substitution_proportion, BatchPlan &plan);
void fill_batch(const BatchPlan &plan, FastxReader *r1, FastxReader *r2, bool keep_comment, int n_threads, ReadBatch &b);

int host_core_share();                       // cores this process may use (all of them, or its share under a one-process-per-GPU launcher)
int host_fill_threads(int n_devices = 1);    // copy threads for fill_batch(): this process's share of the cores
int host_parse_threads(int n_devices = 1);   // FASTQ parser threads per input file

// Fills `b` with the next batch. Returns false when no read could be read (end of input).
bool read_batch(int64_t chunk_size, FastxReader *r1, FastxReader *r2, bool keep_comment, int undirectional,
                float substitution_proportion, ReadBatch &b);

} // namespace bsb
