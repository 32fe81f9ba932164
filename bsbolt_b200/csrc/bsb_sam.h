// bsb_sam.h -- SAM record text (mem_aln2sam, bwamem.c:829-1053, with BSBolt's strand/conversion tags), written once
// for every place that needs it: the host formatter (std::string sink), and the device formatter kernels
// (k_sam_count / k_sam_write in bsb_cuda.cu), which size every record with a counting sink, scan the sizes and then
// write all records of a batch into one text buffer that the host only has to copy to its output.
//
// Everything the text depends on is reached through SamView: plain pointers into the batch arrays, the record arena
// and a flattened contig table, so the same bytes serve host and device.
#pragma once
#include <stddef.h>
#include "bsb_hd.h"

namespace bsb {

struct SamStats { int32_t alignment_score, mapped, bs_conflict, crick, paired; };

struct SamView {
    // options
    int flag;                        // Opt.flag
    int ch_conversion_threshold; float ch_conversion_proportion;
    const char *rg_id; int rg_len;
    // batch
    const char *names; const uint32_t *name_off;
    const char *bases; const char *qual; const uint32_t *seq_off;
    const uint8_t *has_qual, *pattern;
    const char *cmt; const uint32_t *cmt_off;       // may be null (no comments kept)
    // contigs
    const char *ctg_text; const uint32_t *ctg_name_off;   // names back to back, n+1 offsets
    const uint32_t *ctg_anno_off;                         // annotation text follows the names in ctg_text (n+1 offsets), may be null
    const uint8_t *ctg_is_crick, *ctg_sign;               // sign bit 0: the name contains '+', bit 1: it contains '-'
    // results
    const uint8_t *arena; const ReadOut *reads;
};

BSB_HD int sam_nt4(unsigned char c)
{   // nst_nt4_table (bntseq.c:48-65)
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

// Sinks: ch(c), mem(p, n), num(v). The counting sink only adds up.
struct SamCount {
    size_t n = 0;
    BSB_HD void ch(char) { ++n; }
    BSB_HD void mem(const char *, size_t l) { n += l; }
    BSB_HD void pa(double) {}
    BSB_HD void num(long v)
    {
        unsigned long x = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
        int l = v < 0 ? 1 : 0;
        do { ++l; x /= 10; } while (x);
        n += (size_t)l;
    }
};
struct SamWrite {
    char *p;
    BSB_HD void ch(char c) { *p++ = c; }
    BSB_HD void mem(const char *s, size_t l) { for (size_t i = 0; i < l; ++i) p[i] = s[i]; p += l; }
    BSB_HD void num(long v)
    {
        char buf[24]; int l = 0; unsigned long x = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
        do { buf[l++] = (char)('0' + x % 10); x /= 10; } while (x);
        if (v < 0) buf[l++] = '-';
        while (l) *p++ = buf[--l];
    }
    BSB_HD void pa(double) {}   // "\tpa:f:%.3f" (ALT contigs) is only produced by the host sink
};

#if defined(__CUDACC__)
// Device sink of k_sam_write. A thread's record is several hundred bytes; written and read a byte at a time every byte is
// its own L1 transaction (ncu: the byte-wise kernel ran at 84 % of the L1 throughput peak with 10 % of the issue slots busy).
// Here output bytes are collected in a 64-bit register and leave as aligned 8-byte stores (the first bytes up to the next
// 8-byte boundary and the last few go out singly: the words they share belong to the neighbouring records), and the bulk
// sources (qualities, read bases, MD/XB strings, names) are read with aligned 8-byte loads.
struct SamWriteDev {
    char *p; uint64_t acc; int fill;     // fill bytes of acc wait for the aligned word at p
    __device__ __forceinline__ void ch(char c)
    {
        if (fill == 0 && (reinterpret_cast<uintptr_t>(p) & 7)) { *p++ = c; return; }
        acc |= (uint64_t)(uint8_t)c << (fill << 3);
        if (++fill == 8) { *reinterpret_cast<uint64_t *>(p) = acc; p += 8; acc = 0; fill = 0; }
    }
    __device__ __forceinline__ void put8(uint64_t w)      // eight bytes, first byte in the low bits
    {
        if (fill == 0 && (reinterpret_cast<uintptr_t>(p) & 7)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) ch((char)(w >> (k << 3)));
            return;
        }
        const int sh = fill << 3;
        *reinterpret_cast<uint64_t *>(p) = fill ? (acc | w << sh) : w;
        p += 8;
        acc = fill ? w >> (64 - sh) : 0;
    }
    __device__ __forceinline__ void mem(const char *s, size_t l)
    {
        size_t i = 0;
        while (i < l && (reinterpret_cast<uintptr_t>(s + i) & 7)) ch(s[i++]);
        for (; i + 8 <= l; i += 8) put8(*reinterpret_cast<const uint64_t *>(s + i));
        for (; i < l; ++i) ch(s[i]);
    }
    __device__ __forceinline__ void num(long v)
    {
        char buf[24]; int l = 0; unsigned long x = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
        do { buf[l++] = (char)('0' + x % 10); x /= 10; } while (x);
        if (v < 0) buf[l++] = '-';
        while (l) ch(buf[--l]);
    }
    __device__ __forceinline__ void pa(double) {}
    __device__ __forceinline__ void finish() { for (int k = 0; k < fill; ++k) p[k] = (char)(acc >> (k << 3)); p += fill; fill = 0; acc = 0; }
};
#endif

template <class S> BSB_HD void sam_lit(S &o, const char *s) { size_t l = 0; while (s[l]) ++l; o.mem(s, l); }

// SEQ and QUAL columns (bwamem.c:936-966). tab: "ACGTN" for the forward strand, "TGCAN" for the reverse complement.
template <class S> BSB_HD void sam_bases_fwd(S &o, const char *bases, int qb, int qe) { for (int i = qb; i < qe; ++i) o.ch("ACGTN"[sam_nt4((unsigned char)bases[i])]); }
template <class S> BSB_HD void sam_bases_rc(S &o, const char *bases, int qb, int qe) { for (int i = qe - 1; i >= qb; --i) o.ch("TGCAN"[sam_nt4((unsigned char)bases[i])]); }
template <class S> BSB_HD void sam_qual_rev(S &o, const char *qual, int qb, int qe) { for (int i = qe - 1; i >= qb; --i) o.ch(qual[i]); }
#if defined(__CUDACC__)
// the same columns with aligned 8-byte loads of the source
__device__ __forceinline__ void sam_bases_fwd(SamWriteDev &o, const char *bases, int qb, int qe)
{
    int i = qb;
    while (i < qe && (reinterpret_cast<uintptr_t>(bases + i) & 7)) { o.ch("ACGTN"[sam_nt4((unsigned char)bases[i])]); ++i; }
    for (; i + 8 <= qe; i += 8) {
        const uint64_t w = *reinterpret_cast<const uint64_t *>(bases + i);
        uint64_t r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r |= (uint64_t)(uint8_t)"ACGTN"[sam_nt4((unsigned char)(w >> (k << 3)))] << (k << 3);
        o.put8(r);
    }
    for (; i < qe; ++i) o.ch("ACGTN"[sam_nt4((unsigned char)bases[i])]);
}
__device__ __forceinline__ void sam_bases_rc(SamWriteDev &o, const char *bases, int qb, int qe)
{
    int i = qe;                                            // bases[qb, i) are still to be written, last one first
    while (i > qb && (reinterpret_cast<uintptr_t>(bases + i) & 7)) { --i; o.ch("TGCAN"[sam_nt4((unsigned char)bases[i])]); }
    for (; i - 8 >= qb; i -= 8) {
        const uint64_t w = *reinterpret_cast<const uint64_t *>(bases + i - 8);
        uint64_t r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r |= (uint64_t)(uint8_t)"TGCAN"[sam_nt4((unsigned char)(w >> ((7 - k) << 3)))] << (k << 3);
        o.put8(r);
    }
    while (i > qb) { --i; o.ch("TGCAN"[sam_nt4((unsigned char)bases[i])]); }
}
__device__ __forceinline__ void sam_qual_rev(SamWriteDev &o, const char *qual, int qb, int qe)
{
    int i = qe;
    while (i > qb && (reinterpret_cast<uintptr_t>(qual + i) & 7)) { --i; o.ch(qual[i]); }
    for (; i - 8 >= qb; i -= 8) o.put8(__byte_perm((uint32_t)(*reinterpret_cast<const uint64_t *>(qual + i - 8) >> 32), 0, 0x0123)
                                      | (uint64_t)__byte_perm((uint32_t)*reinterpret_cast<const uint64_t *>(qual + i - 8), 0, 0x0123) << 32);
    while (i > qb) { --i; o.ch(qual[i]); }
}
#endif

template <class S>
BSB_HD void sam_cigar(S &o, int n, const uint32_t *cig, const char *ops, int clip_as)
{
    for (int i = 0; i < n; ++i) {
        int c = cig[i] & 0xf;
        if (clip_as >= 0 && (c == 3 || c == 4)) c = clip_as;
        o.num((long)(cig[i] >> 4));
        o.ch(ops[c]);
    }
}

BSB_HD int sam_rlen(int n_cigar, const uint32_t *cigar)
{
    int l = 0;
    for (int k = 0; k < n_cigar; ++k) { int op = cigar[k] & 0xf; if (op == 0 || op == 2) l += cigar[k] >> 4; }
    return l;
}

template <class S> BSB_HD void sam_ctg(S &o, const SamView &v, int rid) { o.mem(v.ctg_text + v.ctg_name_off[rid], v.ctg_name_off[rid + 1] - v.ctg_name_off[rid]); }

// text built by mem_gen_alt (bwamem_extra.c:126-145)
template <class S>
BSB_HD void sam_xa(S &o, const SamView &v, const AlnOut &p)
{
    const XaOut *xa = reinterpret_cast<const XaOut *>(v.arena + p.xa_off);
    for (int i = 0; i < p.xa_n; ++i) {
        const XaOut &t = xa[i];
        sam_ctg(o, v, t.rid);
        o.ch(',');
        o.ch("+-"[t.is_rev]);
        o.num((long)(t.pos + 1));
        o.ch(',');
        sam_cigar(o, t.n_cigar, reinterpret_cast<const uint32_t *>(v.arena + t.cigar_off), "MIDSHN", -1);
        o.ch(',');
        o.num(t.NM);
        if (v.flag & F_XB) { o.ch(','); o.num(t.score); }
        o.ch(';');
    }
}

// countAlts (bs_helpers.cpp:9-16): does the XA text contain the strand character? (the text = contig names, signs,
// positions, CIGARs and numbers: only the names and the sign characters can hold a '+' or '-')
BSB_HD int sam_count_alts(const SamView &v, const AlnOut &p, int is_crick)
{
    const XaOut *xa = reinterpret_cast<const XaOut *>(v.arena + p.xa_off);
    const int want = is_crick ? 0 : 1;      // '+' when the record is on a crick contig, '-' otherwise
    for (int i = 0; i < p.xa_n; ++i) {
        if ((xa[i].is_rev ? 1 : 0) == want) return 1;
        if (v.ctg_sign[xa[i].rid] & (1 << want)) return 1;
    }
    return 0;
}

struct SamMate { bool present; int64_t pos; int rid, is_rev, n_cigar, rlen, ch_meth, ch_unmeth; };

// One SAM line: record `which` of the `n` records of entry `ei`. Appends to o and accumulates the entry's statistics.
template <class S>
BSB_HD void sam_record(S &o, const SamView &v, int ei, const AlnOut *list, int n, int which, const SamMate &mate, SamStats &st)
{
    AlnOut p = list[which];
    SamMate m = mate;
    const uint8_t *arena = v.arena;
    const uint32_t *cigar = reinterpret_cast<const uint32_t *>(arena + p.cigar_off);
    int is_mate_crick = 0, is_crick = 0, bs_conflict = 0, reverse = 0;
    double ch_meth = 0, ch_unmeth = 0;
    if (p.rid >= 0) {
        is_crick = v.ctg_is_crick[p.rid];
        reverse = p.is_rev ? 1 : 0;
        ch_meth += p.ch_meth; ch_unmeth += p.ch_unmeth;
        if (p.xa_n > 0) bs_conflict = sam_count_alts(v, p, is_crick);
        if (is_crick) {
            p.is_rev = 1;
            if (m.present && m.rid >= 0) {
                ch_unmeth += m.ch_unmeth; ch_meth += m.ch_meth;
                is_mate_crick = v.ctg_is_crick[m.rid];
                m.is_rev = is_mate_crick ? 1 : 0;
                if (is_crick != is_mate_crick) bs_conflict = 1;
            }
        } else {
            p.is_rev = 0;
            if (m.present && m.rid >= 0) m.is_rev = is_mate_crick ? 1 : 0;
        }
    }
    if (bs_conflict) { p.score = 0; p.mapq = 0; }
    p.flag |= p.rid < 0 ? 0x4 : 0;
    p.flag |= m.present && m.rid < 0 ? 0x8 : 0;
    if (p.rid < 0 && m.present && m.rid >= 0) { p.rid = m.rid; p.pos = m.pos; p.n_cigar = 0; }
    if (m.present && m.rid < 0 && p.rid >= 0) { m.rid = p.rid; m.pos = p.pos; m.n_cigar = 0; m.rlen = 0; }
    p.flag |= p.is_rev ? 0x10 : 0;
    p.flag |= m.present && m.is_rev ? 0x20 : 0;

    const int l_seq = (int)(v.seq_off[ei + 1] - v.seq_off[ei]);
    const char *bases = v.bases + v.seq_off[ei];
    const char *qual = v.qual + v.seq_off[ei];
    const bool has_qual = v.has_qual[ei] != 0;
    o.mem(v.names + v.name_off[ei], v.name_off[ei + 1] - v.name_off[ei]);
    o.ch('\t');
    o.num((long)((p.flag & 0xffff) | (p.flag & 0x10000 ? 0x100 : 0)));
    o.ch('\t');
    const bool hard = !(v.flag & F_SOFTCLIP) && !p.is_alt;
    if (p.rid >= 0) {
        sam_ctg(o, v, p.rid); o.ch('\t');
        o.num((long)(p.pos + 1)); o.ch('\t');
        o.num(p.mapq); o.ch('\t');
        if (p.n_cigar) sam_cigar(o, p.n_cigar, cigar, "MIDSH", hard ? (which ? 4 : 3) : -1);
        else o.ch('*');
    } else sam_lit(o, "*\t0\t0\t*");
    o.ch('\t');
    if (m.present && m.rid >= 0) {
        if (p.rid == m.rid) o.ch('=');
        else sam_ctg(o, v, m.rid);
        o.ch('\t');
        o.num((long)(m.pos + 1)); o.ch('\t');
        if (p.rid == m.rid) {
            int64_t p0 = p.pos, p1 = m.pos;
            if (p0 > p1) p0 += sam_rlen(p.n_cigar, cigar) - 1;
            else p1 += m.rlen - 1;
            if (m.n_cigar == 0 || p.n_cigar == 0) o.ch('0');
            else o.num((long)(-(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0))));
        } else o.ch('0');
    } else sam_lit(o, "*\t0\t0");
    o.ch('\t');
    if (p.flag & 0x100) {
        sam_lit(o, "*\t*");
    } else {
        int qb = 0, qe = l_seq;
        if (p.n_cigar && which && hard) {
            int c0 = cigar[0] & 0xf, c1 = cigar[p.n_cigar - 1] & 0xf;
            if (!reverse) {
                if (c0 == 4 || c0 == 3) qb += cigar[0] >> 4;
                if (c1 == 4 || c1 == 3) qe -= cigar[p.n_cigar - 1] >> 4;
            } else {
                if (c0 == 4 || c0 == 3) qe -= cigar[0] >> 4;
                if (c1 == 4 || c1 == 3) qb += cigar[p.n_cigar - 1] >> 4;
            }
        }
        if (!reverse) {
            sam_bases_fwd(o, bases, qb, qe);   // '-' (code 5) prints the NUL, like the reference
            o.ch('\t');
            if (has_qual) o.mem(qual + qb, (size_t)(qe > qb ? qe - qb : 0));
            else o.ch('*');
        } else {
            sam_bases_rc(o, bases, qb, qe);
            o.ch('\t');
            if (has_qual) sam_qual_rev(o, qual, qb, qe);
            else o.ch('*');
        }
    }
    if (p.n_cigar) {
        sam_lit(o, "\tNM:i:"); o.num(p.NM);
        sam_lit(o, "\tMD:Z:"); o.mem(reinterpret_cast<const char *>(arena + p.md_off), (size_t)p.md_len);
        if (ch_unmeth + ch_meth >= v.ch_conversion_threshold) {
            double prop = ch_meth / (ch_unmeth + ch_meth);
            sam_lit(o, "\tXC:i:");
            o.ch(prop < v.ch_conversion_proportion ? '0' : '1');
        }
    }
    if (p.score >= 0) { sam_lit(o, "\tAS:i:"); o.num(p.score); }
    if (p.sub >= 0) { sam_lit(o, "\tXS:i:"); o.num(p.sub); }
    if (v.rg_len > 0) { sam_lit(o, "\tRG:Z:"); o.mem(v.rg_id, (size_t)v.rg_len); }
    if (p.rid >= 0) {
        const int pattern = v.pattern[ei];
        sam_lit(o, "\tYS:Z:");
        if (is_crick) {
            if (pattern) { sam_lit(o, "C_G2A"); st.mapped = 1; }
            else { sam_lit(o, "C_C2T"); st.mapped = 2; }
            sam_lit(o, "\tXG:Z:GA");
        } else {
            if (pattern) { sam_lit(o, "W_G2A"); st.mapped = 3; }
            else { sam_lit(o, "W_C2T"); st.mapped = 4; }
            sam_lit(o, "\tXG:Z:CT");
        }
    }
    if (bs_conflict) sam_lit(o, "\tYC:i:1");
    if (!(p.flag & 0x100)) {
        int i;
        for (i = 0; i < n; ++i)
            if (i != which && !(list[i].flag & 0x100)) break;
        if (i < n) {
            sam_lit(o, "\tSA:Z:");
            for (i = 0; i < n; ++i) {
                const AlnOut &r = list[i];
                if (i == which || (r.flag & 0x100)) continue;
                sam_ctg(o, v, r.rid); o.ch(',');
                o.num((long)(r.pos + 1)); o.ch(',');
                o.ch("+-"[r.is_rev]); o.ch(',');
                sam_cigar(o, r.n_cigar, reinterpret_cast<const uint32_t *>(arena + r.cigar_off), "MIDSH", -1);
                o.ch(','); o.num(r.mapq);
                o.ch(','); o.num(r.NM);
                o.ch(';');
            }
        }
        if (p.alt_sc > 0) o.pa((double)p.score / p.alt_sc);   // needs printf("%.3f"): host sink only (the device path is off when ALT contigs exist)
    }
    if (p.xa_n > 0) {
        sam_lit(o, (v.flag & F_XB) ? "\tXB:Z:" : "\tXA:Z:");
        sam_xa(o, v, p);
    }
    if (v.cmt && v.cmt_off[ei + 1] > v.cmt_off[ei]) {
        o.ch('\t');
        o.mem(v.cmt + v.cmt_off[ei], v.cmt_off[ei + 1] - v.cmt_off[ei]);
    }
    if ((v.flag & F_REF_HDR) && p.rid >= 0 && v.ctg_anno_off && v.ctg_anno_off[p.rid + 1] > v.ctg_anno_off[p.rid]) {
        sam_lit(o, "\tXR:Z:");
        for (uint32_t k = v.ctg_anno_off[p.rid]; k < v.ctg_anno_off[p.rid + 1]; ++k) o.ch(v.ctg_text[k] == '\t' ? ' ' : v.ctg_text[k]);
    }
    st.alignment_score += p.score;
    st.bs_conflict = bs_conflict;
    st.crick = is_crick;
    if (m.present) st.paired = 1;
    o.ch('\n');
}

// all records of one entry
template <class S>
BSB_HD void sam_entry(S &o, const SamView &v, int ei, bool is_pe, SamStats &st)
{
    const ReadOut &ro = v.reads[ei];
    const AlnOut *list = reinterpret_cast<const AlnOut *>(v.arena + ro.aln_off);
    SamMate mv;
    mv.present = is_pe;
    mv.pos = ro.h_pos; mv.rid = ro.h_rid; mv.is_rev = ro.h_is_rev; mv.n_cigar = ro.h_n_cigar; mv.rlen = ro.h_rlen;
    mv.ch_meth = ro.h_ch_meth; mv.ch_unmeth = ro.h_ch_unmeth;
    st.alignment_score = st.mapped = st.bs_conflict = st.crick = st.paired = 0;
    for (int k = 0; k < ro.n_aln; ++k) sam_record(o, v, ei, list, ro.n_aln, k, mv, st);
}

} // namespace bsb
