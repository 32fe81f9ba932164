// bsb_bam.h -- SAM record text -> BAM record bytes, and the read-group arbiter, written once for host and device.
//
// The reference's default output is `bwa mem ... | stream_bam` (bsbolt/Align/AlignReads.py:51-60,
// bsbolt/External/HTSLIB/stream_bam.c): htslib parses every SAM line (sam_parse1, sam.c:1924-2160) and writes it as a BAM
// record (bam_write1, sam.c:661-735). bam_record() below is that pair of functions for the lines this aligner prints: the
// *uncompressed* BAM stream is byte-identical to the reference's (tests/test_bam_output.py). It is a template over a byte
// sink (the SAM sinks of bsb_sam.h: counting, plain, and the 8-byte-store device sink) and over the contig lookup, so the
// host writer (host_bam.cpp) and the device kernels (k_bam_count / k_bam_write, bsb_cuda.cu) run the same statements.
//
// bam_arbitrate() is samSorter (bs_sorter.cpp:84-170) for one read-name group: the decision the host makes in
// sam_sort_range (host_mem.cpp), taken on the device when the records never visit the host as text.
#pragma once
#include <stddef.h>
#include <stdlib.h>
#include "bsb_hd.h"
#include "bsb_sam.h"

namespace bsb {

enum : int {
    BAM_OK = 0, BAM_E_TRUNC, BAM_E_QNAME, BAM_E_FLAG, BAM_E_POS, BAM_E_MAPQ, BAM_E_CIGAR_OP, BAM_E_NO_CIGAR, BAM_E_MANY_CIGAR,
    BAM_E_PNEXT, BAM_E_TLEN, BAM_E_CIGAR_SEQ, BAM_E_SEQ_QUAL, BAM_E_QUAL_CHAR, BAM_E_AUX, BAM_E_AUX_TYPE, BAM_E_AUX_FLOAT
};

inline const char *bam_strerror(int rc)
{
    static const char *const msg[] = {
        "ok", "truncated record", "query name too long", "malformed FLAG", "malformed POS", "malformed MAPQ",
        "unrecognized CIGAR operator", "no CIGAR operations", "more than 65535 CIGAR operations", "malformed PNEXT", "malformed TLEN",
        "CIGAR and query sequence are of different length", "SEQ and QUAL are of different length", "invalid QUAL character",
        "incomplete aux field", "unsupported aux type", "float aux field (host encoder only)"};
    return rc >= 0 && rc <= BAM_E_AUX_FLOAT ? msg[rc] : "unknown error";
}

// the fixed 36 bytes in front of every record (SAM spec 4.2; bam_write1 writes block_size + the 32-byte core)
struct BamCore {
    int32_t block_size, tid, pos;
    uint32_t bin_mq_nl, flag_nc, l_seq;
    int32_t mtid, mpos, isize;
};

// seq_nt16_table (hts.c:73-91): "=ACMGRSVTWYHKDBN" in either case, the digits 0-3 as A C G T, anything else N (15)
BSB_HD int bam_nt16(unsigned char c)
{
    switch (c) {
        case '=': return 0;
        case 'A': case 'a': case '0': return 1;
        case 'C': case 'c': case '1': return 2;
        case 'M': case 'm': return 3;
        case 'G': case 'g': case '2': return 4;
        case 'R': case 'r': return 5;
        case 'S': case 's': return 6;
        case 'V': case 'v': return 7;
        case 'T': case 't': case '3': return 8;
        case 'W': case 'w': return 9;
        case 'Y': case 'y': return 10;
        case 'H': case 'h': return 11;
        case 'K': case 'k': return 12;
        case 'D': case 'd': return 13;
        case 'B': case 'b': return 14;
        default: return 15;
    }
}

BSB_HD int bam_cigar_op(char c)
{
    switch (c) {
        case 'M': return 0; case 'I': return 1; case 'D': return 2; case 'N': return 3; case 'S': return 4;
        case 'H': return 5; case 'P': return 6; case '=': return 7; case 'X': return 8; case 'B': return 9;
    }
    return -1;
}

// hts_reg2bin(beg, end, 14, 5) (htslib/hts.h:1322-1328)
BSB_HD int bam_reg2bin(int64_t beg, int64_t end)
{
    int l, s = 14, t = ((1 << 15) - 1) / 7;
    for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << ((l << 1) + l))
        if (beg >> s == end >> s) return t + (int)(beg >> s);
    return 0;
}

BSB_HD uint64_t bam_uint(const char *&p, const char *e)
{
    uint64_t v = 0;
    if (p < e && *p == '+') ++p;
    while (p < e && *p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return v;
}
BSB_HD int64_t bam_int(const char *&p, const char *e)
{
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) neg = *p++ == '-';
    const uint64_t v = bam_uint(p, e);
    return neg ? -(int64_t)v : (int64_t)v;
}
BSB_HD const char *bam_find(const char *p, const char *e, char c) { while (p < e && *p != c) ++p; return p < e ? p : nullptr; }

template <class S> BSB_HD void bam_u32(S &o, uint32_t v) { o.ch((char)v); o.ch((char)(v >> 8)); o.ch((char)(v >> 16)); o.ch((char)(v >> 24)); }
template <class S> BSB_HD void bam_core(S &o, const BamCore &c)
{
    bam_u32(o, (uint32_t)c.block_size); bam_u32(o, (uint32_t)c.tid); bam_u32(o, (uint32_t)c.pos); bam_u32(o, c.bin_mq_nl); bam_u32(o, c.flag_nc);
    bam_u32(o, c.l_seq); bam_u32(o, (uint32_t)c.mtid); bam_u32(o, (uint32_t)c.mpos); bam_u32(o, (uint32_t)c.isize);
}

// Contig lookup of the device (and of anything else that holds the flattened contig table of SamView): binary search over
// the contig ids sorted by name (ties by id, so a duplicated name resolves to its first contig like the header's map does).
struct BamContigs {
    const char *text; const uint32_t *name_off; const int32_t *sorted; int n;
    BSB_HD int cmp(int rid, const char *s, size_t l) const            // name(rid) <=> s
    {
        const char *a = text + name_off[rid];
        const size_t la = name_off[rid + 1] - name_off[rid];
        const size_t m = la < l ? la : l;
        for (size_t k = 0; k < m; ++k)
            if (a[k] != s[k]) return (unsigned char)a[k] < (unsigned char)s[k] ? -1 : 1;
        return la < l ? -1 : la > l ? 1 : 0;
    }
    BSB_HD int operator()(const char *s, size_t l) const
    {
        int lo = 0, hi = n;                           // first sorted slot whose name is >= s
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cmp(sorted[mid], s, l) < 0) lo = mid + 1; else hi = mid; }
        return lo < n && cmp(sorted[lo], s, l) == 0 ? sorted[lo] : -1;
    }
};

// One SAM line [p, e) (no newline) as one BAM record.
//   in  == nullptr: the record is sized and its core computed (into *out); nothing is written for the 36 leading bytes
//                   except that the sink is advanced by them (a counting sink).
//   in  != nullptr: the bytes are written, the 36 leading bytes from *in (the result of a sizing call on the same line).
// Returns BAM_OK or the first defect found (the checks and their order follow sam_parse1 for the fields this aligner prints).
template <class S, class T>
BSB_HD int bam_record(S &o, const T &tid_of, const char *p, const char *e, const BamCore *in, BamCore *out)
{
    size_t bytes = 36;                                  // counted beside the sink: block_size
    if (in) bam_core(o, *in); else o.mem(nullptr, 36);
    const char *t = bam_find(p, e, '\t');
    if (!t) return BAM_E_TRUNC;
    const char *qn = p; const size_t l_qn = (size_t)(t - p); p = t + 1;
    if (l_qn + 1 > 255) return BAM_E_QNAME;
    uint32_t flag = (uint32_t)bam_uint(p, e);
    if (p >= e || *p++ != '\t') return BAM_E_FLAG;
    t = bam_find(p, e, '\t');
    if (!t) return BAM_E_TRUNC;
    int32_t tid = (t - p == 1 && *p == '*') ? -1 : tid_of(p, (size_t)(t - p));
    p = t + 1;
    int64_t pos = (int64_t)bam_uint(p, e) - 1;
    if (p >= e || *p++ != '\t') return BAM_E_POS;
    if (pos < 0 && tid >= 0) tid = -1;
    if (tid < 0) flag |= 4;
    const uint32_t mapq = (uint32_t)bam_uint(p, e) & 0xff;
    if (p >= e || *p++ != '\t') return BAM_E_MAPQ;

    o.mem(qn, l_qn); o.ch(0);
    bytes += l_qn + 1;

    uint32_t n_cigar = 0;
    int64_t cigreflen = 1, qlen_cigar = 0;
    if (p < e && *p != '*') {
        int64_t rlen = 0;
        while (p < e && *p != '\t') {
            const uint64_t len = bam_uint(p, e);
            const int op = p < e ? bam_cigar_op(*p) : -1;
            if (op < 0) return BAM_E_CIGAR_OP;
            ++p;
            bam_u32(o, (uint32_t)(len << 4) | (uint32_t)op);
            ++n_cigar;
            if ((0x3C1A7 >> (op << 1)) & 2) rlen += (int64_t)len;      // bam_cigar_type: bit 1 consumes the reference,
            if ((0x3C1A7 >> (op << 1)) & 1) qlen_cigar += (int64_t)len; //                 bit 0 the query
        }
        if (p >= e || *p++ != '\t') return BAM_E_TRUNC;
        if (n_cigar == 0) return BAM_E_NO_CIGAR;
        if (n_cigar > 0xffff) return BAM_E_MANY_CIGAR;
        bytes += (size_t)n_cigar * 4;
        cigreflen = !(flag & 4) ? rlen : 1;
    } else {
        flag |= 4;
        t = bam_find(p, e, '\t');
        if (!t) return BAM_E_TRUNC;
        p = t + 1;
    }
    const int bin = bam_reg2bin(pos, pos + cigreflen);
    t = bam_find(p, e, '\t');
    if (!t) return BAM_E_TRUNC;
    int32_t mtid;
    if (t - p == 1 && *p == '=') mtid = tid;
    else if (t - p == 1 && *p == '*') mtid = -1;
    else mtid = tid_of(p, (size_t)(t - p));
    p = t + 1;
    int64_t mpos = (int64_t)bam_uint(p, e) - 1;
    if (p >= e || *p++ != '\t') return BAM_E_PNEXT;
    if (mpos < 0 && mtid >= 0) mtid = -1;
    const int64_t isize = bam_int(p, e);
    if (p >= e || *p++ != '\t') return BAM_E_TLEN;
    t = bam_find(p, e, '\t');
    if (!t) return BAM_E_TRUNC;
    uint32_t l_seq = 0;
    if (!(t - p == 1 && *p == '*')) {
        l_seq = (uint32_t)(t - p);
        if (n_cigar && qlen_cigar != (int64_t)l_seq) return BAM_E_CIGAR_SEQ;
        uint32_t i = 0;
        for (; i + 1 < l_seq; i += 2) o.ch((char)(bam_nt16((unsigned char)p[i]) << 4 | bam_nt16((unsigned char)p[i + 1])));
        if (i < l_seq) o.ch((char)(bam_nt16((unsigned char)p[i]) << 4));
        bytes += (l_seq + 1) / 2;
    }
    p = t + 1;
    {   // QUAL: the last mandatory field, ends at a tab or at the end of the line
        t = bam_find(p, e, '\t');
        const char *qe = t ? t : e;
        if (qe - p == 1 && *p == '*') { for (uint32_t i = 0; i < l_seq; ++i) o.ch((char)0xff); }
        else {
            if ((size_t)(qe - p) != l_seq) return BAM_E_SEQ_QUAL;
            for (uint32_t i = 0; i < l_seq; ++i) {
                const int v = (unsigned char)p[i] - 33;
                if (v < 0 || v > 127) return BAM_E_QUAL_CHAR;
                o.ch((char)v);
            }
        }
        bytes += l_seq;
        p = t ? t + 1 : e;
    }
    while (p < e) {   // optional fields TAG:TYPE:VALUE (the types sam_parse1 accepts; 'B' arrays are never printed here)
        t = bam_find(p, e, '\t');
        const char *fe = t ? t : e;
        if (fe - p < 5 || p[2] != ':' || p[4] != ':') return BAM_E_AUX;
        const char type = p[3];
        const char *v = p + 5;
        o.ch(p[0]); o.ch(p[1]);
        bytes += 2;
        if (type == 'A' || type == 'a' || type == 'c' || type == 'C') {
            if (v >= fe) return BAM_E_AUX;
            o.ch('A'); o.ch(*v);
            bytes += 2;
        } else if (type == 'i' || type == 'I') {
            if (v >= fe) return BAM_E_AUX;
            if (*v == '-') {            // the smallest type that holds the value (sam_parse1)
                const int64_t x = bam_int(v, fe);
                if (x >= -128) { o.ch('c'); o.ch((char)x); bytes += 2; }
                else if (x >= -32768) { o.ch('s'); o.ch((char)x); o.ch((char)(x >> 8)); bytes += 3; }
                else { o.ch('i'); bam_u32(o, (uint32_t)(int32_t)x); bytes += 5; }
            } else {
                const uint64_t x = bam_uint(v, fe);
                if (x <= 255) { o.ch('C'); o.ch((char)x); bytes += 2; }
                else if (x <= 65535) { o.ch('S'); o.ch((char)x); o.ch((char)(x >> 8)); bytes += 3; }
                else { o.ch('I'); bam_u32(o, (uint32_t)x); bytes += 5; }
            }
        } else if (type == 'f') {
#if defined(__CUDA_ARCH__)
            return BAM_E_AUX_FLOAT;     // needs strtod: the device path is not taken when such a tag can be printed ("pa:f", ALT contigs)
#else
            char tmp[64];
            size_t l = (size_t)(fe - v) < sizeof tmp - 1 ? (size_t)(fe - v) : sizeof tmp - 1;
            for (size_t k = 0; k < l; ++k) tmp[k] = v[k];
            tmp[l] = 0;
            const float f = (float)strtod(tmp, nullptr);
            uint32_t u;
            const unsigned char *fb = reinterpret_cast<const unsigned char *>(&f);
            u = (uint32_t)fb[0] | (uint32_t)fb[1] << 8 | (uint32_t)fb[2] << 16 | (uint32_t)fb[3] << 24;
            o.ch('f'); bam_u32(o, u);
            bytes += 5;
#endif
        } else if (type == 'Z' || type == 'H') {
            o.ch(type);
            o.mem(v, (size_t)(fe - v));
            o.ch(0);
            bytes += (size_t)(fe - v) + 2;
        } else return BAM_E_AUX_TYPE;
        p = t ? t + 1 : e;
    }
    if (out) {
        out->block_size = (int32_t)(bytes - 4);
        out->tid = tid; out->pos = (int32_t)pos;
        out->bin_mq_nl = (uint32_t)bin << 16 | mapq << 8 | (uint32_t)(l_qn + 1);
        out->flag_nc = flag << 16 | (n_cigar & 0xffff);
        out->l_seq = l_seq;
        out->mtid = mtid; out->mpos = (int32_t)mpos; out->isize = (int32_t)isize;
    }
    return BAM_OK;
}

// The record samSorter::setUnmapped (bs_sorter.cpp:51-82) prints for an entry of a BS-ambiguous read, as BAM bytes: what
// bam_record makes of "<name>\t<77|141|4>\t*\t0\t0\t*\t*\t0\t0\t<SEQ>\t<QUAL>\tAS:i:0\tYS:Z:WC" (set_unmapped, host_mem.cpp).
// qual == nullptr: the read came without qualities -- the text then has an EMPTY quality column, which sam_parse1 rejects
// unless the read is empty too.
template <class S>
BSB_HD int bam_unmapped(S &o, const char *name, size_t l_name, int paired, int first, const char *bases, uint32_t l_seq, const char *qual)
{
    if (l_name + 1 > 255) return BAM_E_QNAME;
    if (!qual && l_seq) return BAM_E_SEQ_QUAL;
    BamCore c;
    const uint32_t flag = paired ? (first ? 77u : 141u) : 4u;
    c.block_size = (int32_t)(32 + l_name + 1 + (l_seq + 1) / 2 + l_seq + 4 + 6);
    c.tid = -1; c.pos = -1;
    c.bin_mq_nl = (uint32_t)bam_reg2bin(-1, 0) << 16 | (uint32_t)(l_name + 1);
    c.flag_nc = flag << 16;
    c.l_seq = l_seq; c.mtid = -1; c.mpos = -1; c.isize = 0;
    bam_core(o, c);
    o.mem(name, l_name); o.ch(0);
    uint32_t i = 0;
    for (; i + 1 < l_seq; i += 2) o.ch((char)(bam_nt16((unsigned char)"ACGTN"[sam_nt4((unsigned char)bases[i])]) << 4 | bam_nt16((unsigned char)"ACGTN"[sam_nt4((unsigned char)bases[i + 1])])));
    if (i < l_seq) o.ch((char)(bam_nt16((unsigned char)"ACGTN"[sam_nt4((unsigned char)bases[i])]) << 4));
    for (i = 0; i < l_seq; ++i) {
        const int v = (unsigned char)qual[i] - 33;
        if (v < 0 || v > 127) return BAM_E_QUAL_CHAR;
        o.ch((char)v);
    }
    o.ch('A'); o.ch('S'); o.ch('C'); o.ch(0);
    o.ch('Y'); o.ch('S'); o.ch('Z'); o.ch('W'); o.ch('C'); o.ch(0);
    return BAM_OK;
}

// ------------------------------------------------------------------------------------------------
// the arbiter, one read-name group at a time
// ------------------------------------------------------------------------------------------------
enum : uint8_t { BAM_DROP = 0, BAM_KEEP = 1, BAM_REWRITE = 2 };   // per entry: not printed / printed as formatted / printed as unmapped

struct MapCounters { unsigned long long reads, alignments, wc2t, wg2a, cc2t, cg2a, unaligned, bs_ambiguous; };

struct ArbiterView {
    const char *names; const uint32_t *name_off;
    const uint8_t *first, *read_group;   // per entry: first mate?; 1 = the second conversion pattern of an undirectional library
    const SamStats *stats; int n;
    BSB_HD bool same_name(int i, int j) const
    {
        const uint32_t li = name_off[i + 1] - name_off[i], lj = name_off[j + 1] - name_off[j];
        if (li != lj) return false;
        const char *a = names + name_off[i], *b = names + name_off[j];
        for (uint32_t k = 0; k < li; ++k) if (a[k] != b[k]) return false;
        return true;
    }
    BSB_HD bool is_head(int i) const { return i == 0 || !same_name(i, i - 1); }
};

// entry `head` starts a group: decide its entries (code[]) and count them (samSorter::processGroup, bs_sorter.cpp:84-150;
// host twin: sam_sort_range, host_mem.cpp)
BSB_HD void bam_arbitrate(const ArbiterView &v, int head, uint8_t *code, MapCounters &c)
{
    long score[2] = {0, 0};
    int end = head;
    for (; end < v.n && (end == head || v.same_name(end, head)); ++end) score[v.read_group[end] ? 1 : 0] += v.stats[end].alignment_score;
    const int pick = score[0] > score[1] ? 0 : score[0] < score[1] ? 1 : 2;
    ++c.reads;
    for (int i = head; i < end; ++i) {
        const int g = v.read_group[i] ? 1 : 0;
        if ((pick == 1) != (g == 1)) { code[i] = BAM_DROP; continue; }      // pick 0 and the tie print group 0, pick 1 prints group 1
        int mapped = v.stats[i].mapped, conflict = v.stats[i].bs_conflict;
        code[i] = BAM_KEEP;
        if (pick == 2) { if (mapped) code[i] = BAM_REWRITE; mapped = 0; conflict = 1; }
        ++c.alignments;
        if (conflict) ++c.bs_ambiguous;
        switch (mapped) {
            case 0: ++c.unaligned; break;
            case 1: ++c.cg2a; break;
            case 2: ++c.cc2t; break;
            case 3: ++c.wg2a; break;
            case 4: ++c.wc2t; break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// all BAM bytes of one entry of a batch whose SAM text is in memory (the device formatter's buffer)
// ------------------------------------------------------------------------------------------------
struct BamView {
    ArbiterView a;
    const char *text; const uint32_t *text_off;      // entry i = text[text_off[i], text_off[i + 1]): its SAM lines, '\n'-terminated
    const uint8_t *code;                             // the arbiter's decision per entry
    BamContigs ctg;
    const char *bases, *qual; const uint32_t *seq_off; const uint8_t *has_qual;
};

// Sizes (S = SamCount) or writes the records of entry i. A written record needs its size in front: the line is parsed twice
// then, first by a counting sink for the 36 leading bytes.
template <class S>
BSB_HD int bam_entry(S &o, const BamView &v, int i, bool writing, int *n_rec = nullptr)
{
    const int what = v.code[i];
    if (what == BAM_DROP) return BAM_OK;
    if (what == BAM_REWRITE) {
        if (n_rec) ++*n_rec;
        const uint32_t l = v.seq_off[i + 1] - v.seq_off[i];
        return bam_unmapped(o, v.a.names + v.a.name_off[i], v.a.name_off[i + 1] - v.a.name_off[i], v.a.stats[i].paired, v.a.first[i],
                            v.bases + v.seq_off[i], l, v.has_qual[i] ? v.qual + v.seq_off[i] : nullptr);
    }
    const char *p = v.text + v.text_off[i], *end = v.text + v.text_off[i + 1];
    while (p < end) {
        const char *nl = bam_find(p, end, '\n');
        const char *e = nl ? nl : end;
        if (e > p) {
            int rc;
            if (writing) {
                SamCount c; BamCore core;
                rc = bam_record(c, v.ctg, p, e, nullptr, &core);
                if (rc == BAM_OK) rc = bam_record(o, v.ctg, p, e, &core, nullptr);
            } else rc = bam_record(o, v.ctg, p, e, nullptr, nullptr);
            if (rc != BAM_OK) return rc;
            if (n_rec) ++*n_rec;
        }
        p = nl ? nl + 1 : end;
    }
    return BAM_OK;
}

// BGZF blocks are cut where an entry starts, like htslib never splits a record that fits a block (bgzf_flush_try, sam.c:728):
// with `quantum` = 0xff00 - (the largest entry of the batch), block k starts at the last entry start <= k * quantum, so that no
// block exceeds 0xff00 bytes and every cut is found on its own (no running fill level). off[0..n]: the entries' offsets.
BSB_HD uint32_t bam_block_cut(const uint32_t *off, int n, uint64_t target)
{
    if (target >= off[n]) return off[n];
    int lo = 0, hi = n;                               // last i with off[i] <= target (off[0] = 0 <= target)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= target) lo = mid; else hi = mid - 1; }
    return off[lo];
}
// the quantum for a batch whose largest entry has max_entry bytes; entries of half a block or more: fixed cuts (records split)
BSB_HD uint32_t bam_block_quantum(uint32_t max_entry) { return max_entry <= 0xff00u / 2 ? 0xff00u - max_entry : 0xff00u; }

} // namespace bsb
