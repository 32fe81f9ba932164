// bsb_final.h -- from alignment regions to SAM-ready records (north_star stages 6, 7, 8).
//
//   mark_primary()     <- mem_mark_primary_se(_core)      (bwamem.c:497-562)
//   approx_mapq()      <- mem_approx_mapq_se              (bwamem.c:1059-1083)
//   bs_md_xb()         <- bwa_gen_cigar2 part 2 + getMethylationContext (bwa.c:180-196, 250-343)
//   aln_head()/align_body()/aln_finish() <- mem_reg2aln + infer_bw (bwamem.c:796-803, 1197-1272), split into the
//                         part decided per region and the base-level part that runs as a deferred AlnTask
//   emit_read()        <- mem_reg2sam + mem_gen_alt       (bwamem.c:1110-1155, bwamem_extra.c:91-152)
//   mate_rescue()      <- mem_matesw                      (bwamem_pair.c:111-180)
//   pair_hits()        <- mem_pair                        (bwamem_pair.c:182-243)
//   finalize_pair()    <- mem_sam_pe                      (bwamem_pair.c:250-395)
//   pestat_candidate() <- first loop of mem_pestat        (bwamem_pair.c:23-60)
//
// libm on the device is not bit-identical to glibc, so every transcendental the reference calls
// (log in MAPQ, log/erfc in pairing) is tabulated by the HOST with glibc for the integer arguments
// that can occur and shipped to HBM (MathTab); the device only does IEEE add/mul/div on them
// (kernels are compiled with -fmad=false so no contraction changes a rounding).
// Attribution: restates BWA-MEM's record selection, MAPQ, pairing and mate rescue (bwamem.c, bwamem_pair.c, bwamem_extra.c;
// GPLv3, Heng Li) and BSBolt's bisulfite tags (bs_helpers.cpp; MIT, Colin P. Farrell). See NOTICE.md.
#pragma once
#include "bsb_extend.h"

namespace bsb {

struct MathTab {
    const double *log_tab;      // log_tab[i] = log((double)i) from glibc, i in [0, n_log)
    int n_log;
    // pairing bonus per orientation: pair_tab[pair_off[d] + (dist - low[d])] =
    //   .721 * log(2. * erfc(fabs((dist - avg)/std) * M_SQRT1_2)) * opt.a      (bwamem_pair.c:218)
    const double *pair_tab;
    int pair_off[4];
};

struct Arena {
    uint8_t *base;
    unsigned long long *used;
    unsigned long long cap;
    BSB_HD uint32_t alloc(uint32_t bytes, int *err)
    {
        unsigned long long need = (bytes + 7ull) & ~7ull;
#if defined(__CUDA_ARCH__)
        unsigned long long off = atomicAdd(used, need);
#else
        unsigned long long off = *used; *used += need;
#endif
        if (off + need > cap) { *err = ERR_ARENA_OVERFLOW; return 0; }
        return (uint32_t)off;
    }
};

struct StrBuf {
    char *s; int n, cap; bool ovf;
    BSB_HD void putc_(char c) { if (n < cap) s[n++] = c; else ovf = true; }
    BSB_HD void putw(int v)
    {
        char buf[16]; int l = 0; unsigned x = v < 0 ? (unsigned)(-v) : (unsigned)v;
        do { buf[l++] = (char)('0' + x % 10); x /= 10; } while (x);
        if (v < 0) buf[l++] = '-';
        while (l) putc_(buf[--l]);
    }
};

struct FinalWS {            // per-thread scratch for the finalisation kernels
    DpScratch dp;
    uint32_t *cigar; int cigar_cap;
    char *md; int md_cap;   // MD + "\tXB:Z:" + XB assembled here
    char *xb; int xb_cap;
    int32_t *cnt;           // [reg_cap] XA bookkeeping
    int8_t *has_alt;        // [reg_cap]
    int32_t *z;             // [reg_cap] mark_primary scratch
    Pair64 *pv, *pu;        // pairing scratch, pair_cap each
    int pair_cap;
    int reg_cap;
    SwScratch sw;
    uint8_t *rev;           // reverse-complemented mate, max_q bytes
};

// ---- primary / secondary marking -----------------------------------------------------------

struct LtRegHash {
    BSB_HD bool operator()(const AlnReg &a, const AlnReg &b) const
    { return a.score > b.score || (a.score == b.score && (a.is_alt < b.is_alt || (a.is_alt == b.is_alt && a.hash < b.hash))); }
};
struct LtRegHash2 {
    BSB_HD bool operator()(const AlnReg &a, const AlnReg &b) const
    { return a.is_alt < b.is_alt || (a.is_alt == b.is_alt && (a.score > b.score || (a.score == b.score && a.hash < b.hash))); }
};

BSB_HD void mark_primary_core(const Opt &opt, int n, AlnReg *a, int32_t *z)
{
    int i, k, tmp, nz = 0;
    tmp = opt.a + opt.b;
    tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
    tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
    z[nz++] = 0;
    for (i = 1; i < n; ++i) {
        for (k = 0; k < nz; ++k) {
            int j = z[k];
            int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * opt.mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i;
        else a[i].secondary = z[k];
    }
}

BSB_HD int mark_primary(const Opt &opt, int n, AlnReg *a, int64_t id, int32_t *z)
{
    int i, n_pri;
    if (n == 0) return 0;
    for (i = n_pri = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash64((uint64_t)(id + i));
        if (!a[i].is_alt) ++n_pri;
    }
    introsort((long)n, a, LtRegHash());
    mark_primary_core(opt, n, a, z);
    for (i = 0; i < n; ++i) {
        AlnReg &p = a[i];
        p.secondary_all = i;
        if (!p.is_alt && p.secondary >= 0 && a[p.secondary].is_alt) p.alt_sc = a[p.secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        if (n_pri > 0) introsort((long)n, a, LtRegHash2());
        for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (a[i].is_alt) a[i].secondary = 0x7fffffff;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {
            for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            mark_primary_core(opt, n_pri, a, z);
        }
    } else {
        for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    return n_pri;
}

// mem_reorder_primary5 (bwamem.c:1085-1108), `bwa mem -5`: the primary with the smallest query start becomes entry 0
BSB_HD void reorder_primary5(int T, int n, AlnReg *a)
{
    int k, n_pri = 0, left_st = 0x7fffffff, left_k = -1;
    for (k = 0; k < n; ++k)
        if (a[k].secondary < 0 && !a[k].is_alt && a[k].score >= T) ++n_pri;
    if (n_pri <= 1) return;
    for (k = 0; k < n; ++k) {
        const AlnReg &p = a[k];
        if (p.secondary >= 0 || p.is_alt || p.score < T) continue;
        if (p.qb < left_st) { left_st = p.qb; left_k = k; }
    }
    if (left_k == 0) return;
    AlnReg t = a[0]; a[0] = a[left_k]; a[left_k] = t;
    for (k = 1; k < n; ++k) {
        AlnReg &p = a[k];
        if (p.secondary == 0) p.secondary = left_k;
        else if (p.secondary == left_k) p.secondary = 0;
        if (p.secondary_all == 0) p.secondary_all = left_k;
        else if (p.secondary_all == left_k) p.secondary_all = 0;
    }
}

// ---- mapping quality -------------------------------------------------------------------------

BSB_HD double tab_log(const MathTab &mt, int x, int *err)
{
    if (x < 0 || x >= mt.n_log) { *err = ERR_SCRATCH_OVERFLOW; return 0.; }
    return mt.log_tab[x];
}

BSB_HD int approx_mapq(const Opt &opt, const MathTab &mt, const AlnReg &a, int *err)
{
    int mapq, l, sub = a.sub ? a.sub : opt.min_seed_len * opt.a;
    double identity;
    sub = a.csub > sub ? a.csub : sub;
    if (sub >= a.score) return 0;
    l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
    identity = 1. - (double)(l * opt.a - a.score) / (opt.a + opt.b) / l;
    if (a.score == 0) {
        mapq = 0;
    } else if (opt.mapQ_coef_len > 0) {
        double tmp;
        tmp = l < opt.mapQ_coef_len ? 1. : opt.mapQ_coef_fac / tab_log(mt, l, err);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a.score - sub) / opt.a * tmp * tmp + .499);
    } else {
        mapq = (int)(30.0 * (1. - (double)sub / a.score) * tab_log(mt, a.seedcov, err) + .499);
        mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
    }
    if (a.sub_n > 0) mapq -= (int)(4.343 * tab_log(mt, a.sub_n + 1, err) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - a.frac_rep) + .499);
    return mapq;
}

BSB_HD int raw_mapq(int diff, int a) { return (int)(6.02 * diff / a + .499); }

// ---- bisulfite-aware MD / XB / NM (stage 8) ----------------------------------------------------

struct MethCounts { int cg_meth, cg_unmeth, ch_meth, ch_unmeth; };

BSB_HD char meth_context(int base1, int base2, int base3, int meth, int *cg, int *ch)
{
    if (base1 == 1 && base2 == 2) { *cg += 1; return meth ? 'X' : 'x'; }
    else if (base1 == 1 && base3 == 2) { *ch += 1; return meth ? 'Y' : 'y'; }
    else { *ch += 1; return meth ? 'Z' : 'z'; }
}

// Walks the CIGAR against the UNCONVERTED reference (opac) and the ORIGINAL read bases.
// `oquery` is the whole read's original bases and, like in the reference (bwamem.c:1228), is NOT
// offset by the clip: x indexes it from 0 (and from l_query-1 downwards on the reverse strand).
// Returns false when the +-2 flank window bridges the strand boundary (no MD is produced there).
BSB_HD bool bs_md_xb(const IndexView &ix, int n_cigar, const uint32_t *cigar, int l_query, const uint8_t *oquery,
                     int64_t rb, int64_t re, StrBuf &md, StrBuf &xb, int *NM, MethCounts *mc)
{
    const int64_t l_pac = ix.l_pac;
    const int reverse = rb > l_pac ? 1 : 0;
    int is_crick = 0;
    if (rb >= ix.crick_l && rb < l_pac) is_crick = 1;
    else if (rb >= ix.crick_l + l_pac) is_crick = 1;
    const int refb = is_crick ? 2 : 1, methb = is_crick ? 0 : 3;
    int64_t cb = rb - 2, ce = re + 2;
    if (ce > l_pac << 1) ce = l_pac << 1;
    if (cb < 0) cb = 0;
    if (!(cb >= l_pac || ce <= l_pac)) return false;
    const int64_t clen = ce - cb;
    if (clen <= 0) return false;
    const bool rev = rb >= l_pac;
    const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
    // cseq'(i): flank-extended unconverted reference in the orientation of the CIGAR. The window lies on one strand, so
    // both orientations walk the forward-strand bytes of opac upwards from one start (complemented for the reverse
    // half: ref_base(p) = 3 - pac[2 l_pac - 1 - p]) -- 32-bit index arithmetic per base instead of 64-bit.
    const int64_t p0 = rev ? (l_pac << 1) - ce : cb;
    const uint8_t *pbase = ix.opac + (p0 >> 2);
    const int poff = (int)(p0 & 3), comp = rev ? 3 : 0, ncl = (int)clen;
    auto cs = [&](int i) -> int {
        if (i < 0 || i >= ncl) return 4;
        const int k = poff + i;
        return comp ^ (pbase[k >> 2] >> ((~k & 3) << 1) & 3);
    };
    auto oq = [&](int x) -> int { return rev ? oquery[l_query - 1 - x] : oquery[x]; };
    int k, x, y, u, n_mm = 0, n_gap = 0, meth_pos = 0;
    mc->cg_meth = mc->cg_unmeth = mc->ch_meth = mc->ch_unmeth = 0;
    for (k = 0, x = y = u = 0; k < n_cigar; ++k) {
        int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
        if (op == 0) {
            for (int i = 0; i < len; ++i) {
                int rc = cs(y + i + 2), qc = oq(x + i);
                if (rc == refb) {
                    int b1, b2, b3;
                    if (!reverse && !is_crick) { b1 = cs(y + i + 2); b2 = cs(y + i + 3); b3 = cs(y + i + 4); }
                    else if (!reverse && is_crick) { b1 = cs(y + i + 1); b2 = cs(y + i + 2); b3 = cs(y + i + 3); }
                    else if (reverse && !is_crick) { b1 = cs(y + i + 2); b2 = cs(y + i + 1); b3 = cs(y + i); }
                    else { b1 = cs(y + i + 3); b2 = cs(y + i + 2); b3 = cs(y + i + 1); }
                    if (qc == refb) {
                        char st = meth_context(b1, b2, b3, 1, &mc->cg_meth, &mc->ch_meth);
                        if (meth_pos > 0) xb.putw(meth_pos);
                        xb.putc_(st);
                        meth_pos = 0; ++u;
                    } else if (qc == methb) {
                        char st = meth_context(b1, b2, b3, 0, &mc->cg_unmeth, &mc->ch_unmeth);
                        if (meth_pos > 0) xb.putw(meth_pos);
                        xb.putc_(st);
                        meth_pos = 0; ++u;
                    } else {
                        md.putw(u); md.putc_(int2base[rc]);
                        ++n_mm; ++meth_pos; u = 0;
                    }
                } else if (qc != rc) {
                    md.putw(u); md.putc_(int2base[rc]);
                    ++n_mm; ++meth_pos; u = 0;
                } else { ++u; ++meth_pos; }
            }
            x += len; y += len;
        } else if (op == 2) {
            if (k > 0 && k < n_cigar - 1) {
                md.putw(u); md.putc_('^');
                for (int i = 0; i < len; ++i) md.putc_(int2base[cs(y + i + 2)]);
                u = 0; n_gap += len;
            }
            y += len;
        } else if (op == 1) { x += len; n_gap += len; meth_pos += len; }
    }
    md.putw(u);
    if (meth_pos > 0) xb.putw(meth_pos);
    *NM = n_mm + n_gap;
    return true;
}

// ---- region -> alignment record ----------------------------------------------------------------

BSB_HD int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
    int w;
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    if (w < iabs(l1 - l2)) w = iabs(l1 - l2);
    return w;
}

struct TaskList {
    AlnTask *a; unsigned int *n; unsigned int cap;
    BSB_HD void push(const AlnTask &t, int *err)
    {
#if defined(__CUDA_ARCH__)
        unsigned int k = atomicAdd(n, 1u);
#else
        unsigned int k = (*n)++;
#endif
        if (k < cap) a[k] = t; else *err = ERR_ARENA_OVERFLOW; // the counter keeps counting: exact size for the retry
    }
};

struct AlnTmp {             // the part of mem_aln_t that does not need the base-level alignment
    int64_t pos; int rid, flag, is_rev, is_alt, mapq, NM, n_cigar, score, sub, alt_sc, md_len;
    MethCounts mc;
};

struct AlnBody {            // the part that does: produced by align_body() / the warp kernel
    int64_t pos; int rid, is_rev, NM, n_cigar, md_len;
    MethCounts mc;
};

BSB_HD void aln_unmapped(AlnTmp &a)
{
    a.pos = -1; a.rid = -1; a.flag = 0x4; a.is_rev = a.is_alt = a.mapq = a.NM = a.n_cigar = 0;
    a.score = a.sub = a.alt_sc = 0; a.md_len = 0;
    a.mc.cg_meth = a.mc.cg_unmeth = a.mc.ch_meth = a.mc.ch_unmeth = 0;
}

// mem_reg2aln, first half: everything that follows from the region record alone
BSB_HD bool aln_head(const Opt &opt, const MathTab &mt, const AlnReg *ar, AlnTmp &a, int *err)
{
    aln_unmapped(a);
    if (ar == nullptr || ar->rb < 0 || ar->re < 0) return false;
    a.flag = 0;
    a.mapq = ar->secondary < 0 ? approx_mapq(opt, mt, *ar, err) : 0;
    if (ar->secondary >= 0) a.flag |= 0x100;
    a.rid = ar->rid;
    a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
    a.is_alt = ar->is_alt; a.alt_sc = ar->alt_sc;
    return true;
}

BSB_HD int band_for_task(const Opt &opt, const AlnTask &t)
{   // infer_bw x2 + clamp (bwamem.c:1219-1223)
    int tmp = infer_bw(t.qe - t.qb, (int)(t.re - t.rb), t.truesc, opt.a, opt.o_del, opt.e_del);
    int w2 = infer_bw(t.qe - t.qb, (int)(t.re - t.rb), t.truesc, opt.a, opt.o_ins, opt.e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > opt.w) w2 = w2 < t.w ? w2 : t.w;
    return w2;
}

// mem_reg2aln after the CIGAR is known: bisulfite MD/XB/NM, strand, position, deletion squeeze, soft clips.
// cig holds the CIGAR of the global alignment (room for two more operations).
BSB_HD void aln_finish(const IndexView &ix, const AlnTask &t, int l_query, const uint8_t *oquery,
                       uint32_t *c, int n_cig, StrBuf &md, StrBuf &xb, AlnBody &b, int *err)
{
    const int qb = t.qb, qe = t.qe;
    const int64_t rb = t.rb, re = t.re;
    int is_rev;
    if (bs_md_xb(ix, n_cig, c, qe - qb, oquery, rb, re, md, xb, &b.NM, &b.mc)) {
        const char tag[7] = "\tXB:Z:";
        for (int j = 0; j < 6; ++j) md.putc_(tag[j]);
        for (int j = 0; j < xb.n; ++j) md.putc_(xb.s[j]);
        if (md.ovf || xb.ovf) *err = ERR_SCRATCH_OVERFLOW;
        b.md_len = md.n;
    } else {
        // The +-2 flank of the unconverted reference bridges the strand boundary: the reference returns
        // the CIGAR without MD/XB and leaves NM = -1, which its 22-bit field prints as 4194303; the MD
        // text it then reads is whatever follows the CIGAR in memory (empty in practice) (bwa.c:266).
        b.NM = 0x3fffff; b.md_len = 0;
        b.mc.cg_meth = b.mc.cg_unmeth = b.mc.ch_meth = b.mc.ch_unmeth = 0;
    }
    int64_t pos = depos(ix.l_pac, rb < ix.l_pac ? rb : re - 1, &is_rev);
    b.is_rev = is_rev;
    int n_cigar = n_cig;
    if (n_cigar > 0) { // squeeze out a leading or trailing deletion
        if ((c[0] & 0xf) == 2) {
            pos += c[0] >> 4;
            --n_cigar;
            for (int j = 0; j < n_cigar; ++j) c[j] = c[j + 1];
        } else if ((c[n_cigar - 1] & 0xf) == 2) --n_cigar;
    }
    if (qb != 0 || qe != l_query) {
        int clip5 = is_rev ? l_query - qe : qb;
        int clip3 = is_rev ? qb : l_query - qe;
        if (clip5) {
            for (int j = n_cigar; j > 0; --j) c[j] = c[j - 1];
            c[0] = (uint32_t)clip5 << 4 | 3;
            ++n_cigar;
        }
        if (clip3) c[n_cigar++] = (uint32_t)clip3 << 4 | 3;
    }
    b.n_cigar = n_cigar;
    b.rid = pos2rid(ix, pos);
    b.pos = pos - ix.anns[b.rid].offset;
}

// mem_reg2aln, second half (scalar form): up to three global alignments with a doubling band, then aln_finish
BSB_HD bool align_body(const Opt &opt, const IndexView &ix, const AlnTask &t, int l_query, const uint8_t *query,
                       const uint8_t *oquery, FinalWS &ws, AlnBody &b, int *err)
{
    int w2 = band_for_task(opt, t), score = 0, last_sc = -(1 << 30), i = 0;
    CigarBuf cig = {ws.cigar, 0, ws.cigar_cap - 2};
    bool ok;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        ok = global_core(opt, ix, w2, t.qe - t.qb, query + t.qb, t.rb, t.re, &score, &cig, ws.dp, err);
        if (!ok) break;
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < t.truesc - opt.a);
    if (!ok) { if (!*err) *err = ERR_NO_MD; return false; }
    StrBuf md = {ws.md, 0, ws.md_cap, false}, xb = {ws.xb, 0, ws.xb_cap, false};
    aln_finish(ix, t, l_query, oquery, ws.cigar, cig.n, md, xb, b, err);
    return true;
}

BSB_HD int get_rlen(int n_cigar, const uint32_t *cigar)
{
    int l = 0;
    for (int k = 0; k < n_cigar; ++k) { int op = cigar[k] & 0xf; if (op == 0 || op == 2) l += cigar[k] >> 4; }
    return l;
}

// Writes a finished body (CIGAR in `cigar`, text in `md`) to its record(s) in the arena.
// lane-0 / single-thread code. `reads` is the ReadOut array of the batch.
BSB_HD void task_store(const IndexView &ix, const AlnTask &t, const AlnBody &b, const uint32_t *cigar, const char *md,
                       Arena &ar, ReadOut *reads, int *err)
{
    uint32_t cig_off = 0, md_off = 0;
    if (b.n_cigar > 0) {
        cig_off = ar.alloc((uint32_t)b.n_cigar * 4u, err);
        if (*err == ERR_ARENA_OVERFLOW) return;
        uint32_t *d = reinterpret_cast<uint32_t *>(ar.base + cig_off);
        for (int j = 0; j < b.n_cigar; ++j) d[j] = cigar[j];
    }
    if (t.kind == 0) {
        if (b.md_len > 0) {
            md_off = ar.alloc((uint32_t)b.md_len, err);
            if (*err == ERR_ARENA_OVERFLOW) return;
            char *d = reinterpret_cast<char *>(ar.base + md_off);
            for (int j = 0; j < b.md_len; ++j) d[j] = md[j];
        }
        AlnOut &o = *reinterpret_cast<AlnOut *>(ar.base + t.target_off);
        o.pos = b.pos; o.rid = b.rid; o.is_rev = b.is_rev; o.NM = b.NM; o.n_cigar = b.n_cigar; o.cigar_off = cig_off;
        o.md_off = md_off; o.md_len = b.md_len;
        o.ch_meth = b.mc.ch_meth; o.ch_unmeth = b.mc.ch_unmeth; o.cg_meth = b.mc.cg_meth; o.cg_unmeth = b.mc.cg_unmeth;
    } else {
        XaOut &x = *reinterpret_cast<XaOut *>(ar.base + t.target_off);
        x.pos = b.pos; x.rid = b.rid; x.NM = b.NM; x.n_cigar = b.n_cigar; x.cigar_off = cig_off;
        x.is_rev = ix.anns[b.rid].is_crick ? 1 : 0;
    }
    if (t.mate_read >= 0) {
        ReadOut &dst = reads[t.mate_read];
        dst.h_pos = b.pos; dst.h_rid = b.rid; dst.h_is_rev = b.is_rev; dst.h_n_cigar = b.n_cigar;
        dst.h_rlen = get_rlen(b.n_cigar, cigar); dst.h_ch_meth = b.mc.ch_meth; dst.h_ch_unmeth = b.mc.ch_unmeth;
    }
}

// header part of a record into the arena; the body fields are completed by the task
BSB_HD void aln_store(const AlnTmp &t, AlnOut &o)
{
    o.pos = t.pos; o.rid = t.rid; o.flag = t.flag; o.is_rev = t.is_rev; o.is_alt = t.is_alt; o.mapq = t.mapq; o.NM = t.NM;
    o.n_cigar = 0; o.md_len = 0;
    o.ch_meth = o.ch_unmeth = o.cg_meth = o.cg_unmeth = 0;
    o.score = t.score; o.sub = t.sub; o.alt_sc = t.alt_sc;
    o.xa_off = 0; o.xa_n = 0; o.cigar_off = 0; o.md_off = 0;
}

BSB_HD void mate_unmapped(ReadOut &dst)
{
    dst.h_pos = -1; dst.h_rid = -1; dst.h_is_rev = 0; dst.h_n_cigar = 0; dst.h_rlen = 0; dst.h_ch_meth = dst.h_ch_unmeth = 0;
}

BSB_HD void push_task(TaskList &tl, const AlnReg &ar, int read, int kind, uint32_t target_off, int mate_read, int *err)
{
    AlnTask t;
    t.rb = ar.rb; t.re = ar.re; t.qb = ar.qb; t.qe = ar.qe; t.truesc = ar.truesc; t.w = ar.w;
    t.read = read; t.kind = kind; t.target_off = target_off; t.mate_read = mate_read;
    tl.push(t, err);
}

// ---- XA + record emission ------------------------------------------------------------------------

BSB_HD int xa_pri_idx(double XA_drop_ratio, const AlnReg *a, int i)
{
    int k = a[i].secondary_all;
    if (k >= 0 && a[i].score >= a[k].score * XA_drop_ratio) return k;
    return -1;
}

struct ReadCtx {           // everything the finalisation of one read needs
    int read;              // bseq entry index
    int l_seq;
    AlnReg *regs; int n_regs;
};

BSB_HD void xa_prepare(const Opt &opt, const ReadCtx &rc, FinalWS &ws)
{
    for (int i = 0; i < rc.n_regs; ++i) { ws.cnt[i] = 0; ws.has_alt[i] = 0; }
    for (int i = 0; i < rc.n_regs; ++i) {
        int r = xa_pri_idx((double)opt.XA_drop_ratio, rc.regs, i);
        if (r >= 0) { ++ws.cnt[r]; if (rc.regs[i].is_alt) ws.has_alt[r] = 1; }
    }
}

// XA entries of region k (mem_gen_alt): one global alignment task per listed secondary hit
BSB_HD void xa_emit(const Opt &opt, const ReadCtx &rc, int k, FinalWS &ws, Arena &ar, TaskList &tl, uint32_t aln_off, int *err)
{
    AlnOut &o = *reinterpret_cast<AlnOut *>(ar.base + aln_off);
    o.xa_n = 0; o.xa_off = 0;
    if (opt.flag & F_ALL) return;
    int cnt = ws.cnt[k];
    if (cnt == 0) return;
    if (cnt > opt.max_XA_hits_alt || (!ws.has_alt[k] && cnt > opt.max_XA_hits)) return;
    uint32_t off = ar.alloc((uint32_t)cnt * (uint32_t)sizeof(XaOut), err);
    if (*err == ERR_ARENA_OVERFLOW) return;
    XaOut *xa = reinterpret_cast<XaOut *>(ar.base + off);
    int n = 0;
    for (int i = 0; i < rc.n_regs; ++i) {
        if (xa_pri_idx((double)opt.XA_drop_ratio, rc.regs, i) != k) continue;
        const AlnReg &r = rc.regs[i];
        XaOut &x = xa[n];
        x.pos = -1; x.rid = r.rid; x.NM = 0; x.score = r.score; x.n_cigar = 0; x.is_rev = 0; x.cigar_off = 0;
        push_task(tl, r, rc.read, 1, off + (uint32_t)n * (uint32_t)sizeof(XaOut), -1, err);
        ++n;
    }
    o.xa_off = off; o.xa_n = n;
}

// mem_reg2sam: choose the regions that become SAM lines; mate_read >= 0: the first record doubles as the
// mate record of that entry (h[] of the no-pairing branch of mem_sam_pe)
BSB_HD void emit_read(const Opt &opt, const MathTab &mt, const ReadCtx &rc, int extra_flag, int mate_read,
                      FinalWS &ws, Arena &ar, TaskList &tl, ReadOut *reads, int *err)
{
    ReadOut &ro = reads[rc.read];
    const AlnReg *a = rc.regs;
    int n_out = 0;
    for (int k = 0; k < rc.n_regs; ++k) {
        const AlnReg &p = a[k];
        if (p.score < opt.T) continue;
        if (p.secondary >= 0 && (p.is_alt || !(opt.flag & F_ALL))) continue;
        if (p.secondary >= 0 && p.secondary < 0x7fffffff && p.score < a[p.secondary].score * opt.drop_ratio) continue;
        ++n_out;
    }
    if (!(opt.flag & F_ALL)) xa_prepare(opt, rc, ws);
    int n_alloc = n_out ? n_out : 1;
    ro.aln_off = ar.alloc((uint32_t)n_alloc * (uint32_t)sizeof(AlnOut), err);
    ro.n_aln = 0;
    if (*err == ERR_ARENA_OVERFLOW) return;
    AlnOut *out = reinterpret_cast<AlnOut *>(ar.base + ro.aln_off);
    if (n_out == 0) {
        AlnTmp t;
        aln_unmapped(t);
        t.flag |= extra_flag;
        aln_store(t, out[0]);
        ro.n_aln = 1;
        if (mate_read >= 0) mate_unmapped(reads[mate_read]);
        return;
    }
    int l = 0, mapq0 = 0;
    for (int k = 0; k < rc.n_regs; ++k) {
        const AlnReg &p = a[k];
        if (p.score < opt.T) continue;
        if (p.secondary >= 0 && (p.is_alt || !(opt.flag & F_ALL))) continue;
        if (p.secondary >= 0 && p.secondary < 0x7fffffff && p.score < a[p.secondary].score * opt.drop_ratio) continue;
        AlnTmp t;
        aln_head(opt, mt, &p, t, err);
        t.flag |= extra_flag;
        if (p.secondary >= 0) t.sub = -1;
        if (l && p.secondary < 0) t.flag |= (opt.flag & F_NO_MULTI) ? 0x10000 : 0x800;
        if (l == 0) mapq0 = t.mapq;
        if (!(opt.flag & F_KEEP_SUPP_MAPQ) && l && !p.is_alt && t.mapq > mapq0) t.mapq = mapq0;
        aln_store(t, out[l]);
        const uint32_t off = ro.aln_off + (uint32_t)l * (uint32_t)sizeof(AlnOut);
        push_task(tl, p, rc.read, 0, off, l == 0 ? mate_read : -1, err);
        xa_emit(opt, rc, k, ws, ar, tl, off, err);
        ++l;
    }
    ro.n_aln = l;
}

// ---- paired-end: insert-size candidates, rescue, pairing ------------------------------------------

BSB_HD int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
    int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
    int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
    *dist = p2 > b1 ? p2 - b1 : b1 - p2;
    return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

BSB_HD int cal_sub(const Opt &opt, const AlnReg *a, int n)
{
    int j;
    for (j = 1; j < n; ++j) {
        int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
        int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
        if (e_min > b_max) {
            int min_l = a[j].qe - a[j].qb < a[0].qe - a[0].qb ? a[j].qe - a[j].qb : a[0].qe - a[0].qb;
            if (e_min - b_max >= min_l * opt.mask_level) break;
        }
    }
    return j < n ? a[j].score : opt.min_seed_len * opt.a;
}

// One pair's contribution to the batch insert-size statistics: returns dir in [0,4) and *is > 0,
// or -1 when the pair is not a confident unique pair (bwamem_pair.c:51-63).
BSB_HD int pestat_candidate(const Opt &opt, int64_t l_pac, const AlnReg *r0, int n0, const AlnReg *r1, int n1, int64_t *is)
{
    if (n0 == 0 || n1 == 0) return -1;
    if (cal_sub(opt, r0, n0) > 0.8 * r0[0].score) return -1;
    if (cal_sub(opt, r1, n1) > 0.8 * r1[0].score) return -1;
    if (r0[0].rid != r1[0].rid) return -1;
    int dir = infer_dir(l_pac, r0[0].rb, r1[0].rb, is);
    if (*is && *is <= opt.max_ins) return dir;
    return -1;
}

// The rescue Smith-Waterman of mem_matesw depends only on the region `a`, the orientation r, the insert-size statistics and
// the mate's sequence -- not on the mate's region list, which only decides whether the orientation is skipped and where the
// result is inserted. So every Smith-Waterman a pair can need is known before the first one runs: they are enumerated as
// RescueJobs, computed side by side (k_rescue_sw, bsb_rescue.h), and the sequential logic is then replayed over the results.
struct RescueJob {
    int64_t rb;          // window start (doubled coordinates); the window is tlen bases long
    int32_t tlen, qlen;  // qlen = length of the mate
    int32_t key;         // side << 16 | j << 2 | r : mem_matesw call (side, j), orientation r
    int32_t mate_read;   // bseq entry of the mate (the query)
    int32_t is_rev;      // the query is the reverse complement of the mate
    int32_t pair;        // position of the pair in the heavy list
};
struct RescueSink {      // enumeration: where the jobs go (jobs == nullptr: they are only counted)
    RescueJob *jobs; unsigned int *n; unsigned int cap;
    int32_t pair, mate_read;
};
struct RescuePre { const RescueJob *jobs; const SwResult *res; int n; };   // replay: the jobs of this pair and their results

// window of orientation r around region a (bwamem_pair.c:129-147); false: no Smith-Waterman for this orientation
BSB_HD bool rescue_window(const Opt &opt, const IndexView &ix, const PeStat pes[4], const AlnReg &a, int l_ms, int r, int64_t *rb_, int64_t *re_, int *is_rev_)
{
    const int64_t l_pac = ix.l_pac;
    const int is_rev = (r >> 1 != (r & 1)), is_larger = !(r >> 1);
    int64_t rb, re;
    if (!is_rev) {
        rb = is_larger ? a.rb + pes[r].low : a.rb - pes[r].high;
        re = (is_larger ? a.rb + pes[r].high : a.rb - pes[r].low) + l_ms;
    } else {
        rb = (is_larger ? a.rb + pes[r].low : a.rb - pes[r].high) - l_ms;
        re = is_larger ? a.rb + pes[r].high : a.rb - pes[r].low;
    }
    if (rb < 0) rb = 0;
    if (re > l_pac << 1) re = l_pac << 1;
    int rid = -1;
    bool have_ref = false;
    if (rb < re) { rid = fetch_window(ix, &rb, (rb + re) >> 1, &re); have_ref = true; }
    *rb_ = rb; *re_ = re; *is_rev_ = is_rev;
    return have_ref && a.rid == rid && re - rb >= opt.min_seed_len;
}

// mem_matesw: Smith-Waterman of the mate inside the window implied by region `a` and the
// insert-size distribution. `ma` is the mate's region list (grows in place).
// `defer` (optional): instead of running the Smith-Waterman, report that one is needed and return before anything has
// been modified -- the pair is then finalised by the rescue kernels.
// `sink` (optional): enumeration only -- every Smith-Waterman this call can need becomes a RescueJob, nothing is modified.
// `pre` (optional): replay -- the Smith-Waterman results are taken from the jobs computed before.
BSB_HD int mate_rescue(const Opt &opt, const IndexView &ix, const PeStat pes[4], const AlnReg &a, int l_ms, const uint8_t *ms,
                       RegList &ma, FinalWS &ws, int *err, int *defer = nullptr, int call_key = 0, RescueSink *sink = nullptr,
                       const RescuePre *pre = nullptr)
{
    const int64_t l_pac = ix.l_pac;
    int i, r, skip[4], n = 0;
    if (sink) {
        for (r = 0; r < 4; ++r) {
            int64_t rb, re; int is_rev;
            if (pes[r].failed || !rescue_window(opt, ix, pes, a, l_ms, r, &rb, &re, &is_rev)) continue;
            const unsigned int k = (*sink->n)++;     // the counter belongs to the enumerating thread
            if (sink->jobs && k < sink->cap) {
                RescueJob jb;
                jb.rb = rb; jb.tlen = (int32_t)(re - rb); jb.qlen = l_ms; jb.key = call_key | r; jb.mate_read = sink->mate_read;
                jb.is_rev = is_rev; jb.pair = sink->pair;
                sink->jobs[k] = jb;
            }
        }
        return 0;
    }
    for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
    for (i = 0; i < ma.n; ++i) {
        int64_t dist;
        r = infer_dir(l_pac, a.rb, ma.a[i].rb, &dist);
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
    }
    if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
    for (r = 0; r < 4; ++r) {
        int is_rev;
        int64_t rb, re;
        if (skip[r]) continue;
        if (rescue_window(opt, ix, pes, a, l_ms, r, &rb, &re, &is_rev)) {
            if (defer) { *defer = 1; return n; }
            SwResult aln;
            if (pre) {
                int k = 0;
                while (k < pre->n && pre->jobs[k].key != (call_key | r)) ++k;
                if (k == pre->n) { *err = ERR_SCRATCH_OVERFLOW; return n; }   // cannot happen: the enumeration saw the same windows
                aln = pre->res[k];
            } else {
                const uint8_t *seq = ms;
                if (is_rev) {
                    for (i = 0; i < l_ms; ++i) ws.rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
                    seq = ws.rev;
                }
                int xtra = SW_XSUBO | SW_XSTART | (l_ms * opt.a < 250 ? SW_XBYTE : 0) | (opt.min_seed_len * opt.a);
                QrySeq q = {seq, 1};
                RefSeq t = {ix.pac, l_pac, rb, 1};
                aln = sw_local(l_ms, q, (int)(re - rb), t, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, xtra, ws.sw, err);
            }
            if (aln.score >= opt.min_seed_len && aln.qb >= 0) {
                AlnReg b;
                alnreg_clear(b);
                b.rid = a.rid;
                b.is_alt = a.is_alt;
                b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
                b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
                b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
                b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
                b.score = aln.score;
                b.csub = aln.score2;
                b.secondary = -1;
                b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
                if (ma.n >= ma.cap) { *err = ERR_SCRATCH_OVERFLOW; return n; }
                ++ma.n;
                for (i = 0; i < ma.n - 1; ++i)
                    if (ma.a[i].score < b.score) break;
                int tmp = i;
                for (i = ma.n - 1; i > tmp; --i) ma.a[i] = ma.a[i - 1];
                ma.a[i] = b;
            }
            ++n;
        }
        if (n) ma.n = sort_dedup_patch(opt, ix, nullptr, ma.n, ma.a, ws.dp, err);
    }
    return n;
}

// mem_pair: best consistent pair among the primary hits of both ends
BSB_HD int pair_hits(const Opt &opt, const IndexView &ix, const MathTab &mt, const PeStat pes[4],
                     const AlnReg *a0, const AlnReg *a1, int id, int *sub, int *n_sub, int z[2], const int n_pri[2],
                     FinalWS &ws, int *err)
{
    Pair64 *v = ws.pv, *u = ws.pu;
    int nv = 0, nu = 0;
    int r, i, k, y[4], ret;
    const int64_t l_pac = ix.l_pac;
    for (r = 0; r < 2; ++r) {
        const AlnReg *a = r ? a1 : a0;
        for (i = 0; i < n_pri[r]; ++i) {
            const AlnReg &e = a[i];
            Pair64 key;
            key.x = (uint64_t)(e.rb < l_pac ? e.rb : (l_pac << 1) - 1 - e.rb);
            key.x = (uint64_t)e.rid << 32 | (key.x - (uint64_t)ix.anns[e.rid].offset);
            key.y = (uint64_t)e.score << 32 | (uint64_t)(int64_t)(i << 2 | (e.rb >= l_pac) << 1 | r);
            if (nv >= ws.pair_cap) { *err = ERR_SCRATCH_OVERFLOW; return 0; }
            v[nv++] = key;
        }
    }
    introsort((long)nv, v, LtPair64());
    y[0] = y[1] = y[2] = y[3] = -1;
    for (i = 0; i < nv; ++i) {
        for (r = 0; r < 2; ++r) {
            int dir = r << 1 | (int)(v[i].y >> 1 & 1), which;
            if (pes[dir].failed) continue;
            which = r << 1 | (int)((v[i].y & 1) ^ 1);
            if (y[which] < 0) continue;
            for (k = y[which]; k >= 0; --k) {
                int64_t dist;
                int q;
                if ((int)(v[k].y & 3) != which) continue;
                dist = (int64_t)v[i].x - (int64_t)v[k].x;
                if (dist > pes[dir].high) break;
                if (dist < pes[dir].low) continue;
                double bonus = mt.pair_tab[mt.pair_off[dir] + (int)(dist - pes[dir].low)];
                q = (int)((double)((v[i].y >> 32) + (v[k].y >> 32)) + bonus + .499);
                if (q < 0) q = 0;
                if (nu >= ws.pair_cap) { *err = ERR_SCRATCH_OVERFLOW; return 0; }
                Pair64 &p = u[nu++];
                p.y = (uint64_t)k << 32 | (uint32_t)i;
                // `id` is an int in the reference and id<<8 wraps in 32 bits before widening
                p.x = (uint64_t)q << 32 | (hash64(p.y ^ (uint64_t)(int64_t)(int32_t)((uint32_t)id << 8)) & 0xffffffffU);
            }
        }
        y[v[i].y & 3] = i;
    }
    if (nu) {
        int tmp = opt.a + opt.b;
        tmp = tmp > opt.o_del + opt.e_del ? tmp : opt.o_del + opt.e_del;
        tmp = tmp > opt.o_ins + opt.e_ins ? tmp : opt.o_ins + opt.e_ins;
        introsort((long)nu, u, LtPair64());
        i = (int)(u[nu - 1].y >> 32); k = (int)(u[nu - 1].y << 32 >> 32);
        z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
        z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
        ret = (int)(u[nu - 1].x >> 32);
        *sub = nu > 1 ? (int)(u[nu - 2].x >> 32) : 0;
        for (i = nu - 2, *n_sub = 0; i >= 0; --i)
            if (*sub - (int)(u[i].x >> 32) <= tmp) ++*n_sub;
    } else { ret = 0; *sub = 0; *n_sub = 0; }
    return ret;
}

// mem_sam_pe. regs0/regs1 are the two ends' region lists (capacity cap0/cap1; rescue may append).
// `first` is the bseq index of read 1 of the pair (read 2 is first+1); alignments are queued in `tl`.
BSB_HD void finalize_pair(const Opt &opt, const IndexView &ix, const MathTab &mt, const PeStat pes[4], uint64_t id, int first,
                          int l0, const uint8_t *seq0, RegList &r0, int l1, const uint8_t *seq1, RegList &r1,
                          FinalWS &ws, Arena &ar, TaskList &tl, ReadOut *reads, int *err, int *defer = nullptr,
                          RescueSink *sink = nullptr, const RescuePre *pre = nullptr)
{
    RegList *a[2] = {&r0, &r1};
    const int ls[2] = {l0, l1};
    const uint8_t *seqs[2] = {seq0, seq1};
    ReadOut *ro[2] = {&reads[first], &reads[first + 1]};
    int i, j, z[2] = {0, 0}, o, subo = 0, n_sub = 0, extra_flag = 1, n_pri[2];

    if (!(opt.flag & F_NO_RESCUE)) {
        // b[i]: copies of the good hits of end i taken BEFORE any rescue modifies the lists
        // (at most max_matesw are used); they live at the tail of the other scratch arrays
        AlnReg *bcopy[2]; int nb[2];
        for (i = 0; i < 2; ++i) {
            nb[i] = 0;
            bcopy[i] = a[i]->a + a[i]->cap; // caller reserves max_matesw slots past cap
            for (j = 0; j < a[i]->n; ++j)
                if (a[i]->a[j].score >= a[i]->a[0].score - opt.pen_unpaired) {
                    if (nb[i] < opt.max_matesw) bcopy[i][nb[i]] = a[i]->a[j];
                    ++nb[i];
                }
            if (nb[i] > opt.max_matesw) nb[i] = opt.max_matesw;
        }
        for (i = 0; i < 2; ++i)
            for (j = 0; j < nb[i]; ++j) {
                if (sink) sink->mate_read = first + !i;
                mate_rescue(opt, ix, pes, bcopy[i][j], ls[!i], seqs[!i], *a[!i], ws, err, defer, i << 16 | j << 2, sink, pre);
                if (defer && *defer) return;
            }
        if (sink) return;   // enumeration: the jobs are out, nothing else is touched
    }
    n_pri[0] = mark_primary(opt, a[0]->n, a[0]->a, (int64_t)(id << 1 | 0), ws.z);
    n_pri[1] = mark_primary(opt, a[1]->n, a[1]->a, (int64_t)(id << 1 | 1), ws.z);
    if (opt.flag & F_PRIMARY5) { reorder_primary5(opt.T, a[0]->n, a[0]->a); reorder_primary5(opt.T, a[1]->n, a[1]->a); }
    ReadCtx rc[2];
    for (i = 0; i < 2; ++i) { rc[i].read = first + i; rc[i].l_seq = ls[i]; rc[i].regs = a[i]->a; rc[i].n_regs = a[i]->n; }

    bool paired = false;
    if (!(opt.flag & F_NOPAIRING) && n_pri[0] && n_pri[1] &&
        (o = pair_hits(opt, ix, mt, pes, a[0]->a, a[1]->a, (int)id, &subo, &n_sub, z, n_pri, ws, err)) > 0) {
        int is_multi[2], q_pe, score_un, q_se[2];
        for (i = 0; i < 2; ++i) {
            for (j = 1; j < n_pri[i]; ++j)
                if (a[i]->a[j].secondary < 0 && a[i]->a[j].score >= opt.T) break;
            is_multi[i] = j < n_pri[i] ? 1 : 0;
        }
        if (!(is_multi[0] || is_multi[1])) {
            paired = true;
            score_un = a[0]->a[0].score + a[1]->a[0].score - opt.pen_unpaired;
            subo = subo > score_un ? subo : score_un;
            q_pe = raw_mapq(o - subo, opt.a);
            if (n_sub > 0) q_pe -= (int)(4.343 * tab_log(mt, n_sub + 1, err) + .499);
            if (q_pe < 0) q_pe = 0;
            if (q_pe > 60) q_pe = 60;
            q_pe = (int)(q_pe * (1. - .5 * (a[0]->a[0].frac_rep + a[1]->a[0].frac_rep)) + .499);
            if (o > score_un) {
                AlnReg *c[2] = {&a[0]->a[z[0]], &a[1]->a[z[1]]};
                for (i = 0; i < 2; ++i) {
                    if (c[i]->secondary >= 0) { c[i]->sub = a[i]->a[c[i]->secondary].score; c[i]->secondary = -2; }
                    q_se[i] = approx_mapq(opt, mt, *c[i], err);
                }
                q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
                q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
                extra_flag |= 2;
                q_se[0] = q_se[0] < raw_mapq(c[0]->score - c[0]->csub, opt.a) ? q_se[0] : raw_mapq(c[0]->score - c[0]->csub, opt.a);
                q_se[1] = q_se[1] < raw_mapq(c[1]->score - c[1]->csub, opt.a) ? q_se[1] : raw_mapq(c[1]->score - c[1]->csub, opt.a);
            } else {
                z[0] = z[1] = 0;
                q_se[0] = approx_mapq(opt, mt, a[0]->a[0], err);
                q_se[1] = approx_mapq(opt, mt, a[1]->a[0], err);
            }
            for (i = 0; i < 2; ++i) {
                int k = a[i]->a[z[i]].secondary_all;
                if (k >= 0 && k < n_pri[i]) {
                    for (j = 0; j < a[i]->n; ++j)
                        if (a[i]->a[j].secondary_all == k || j == k) a[i]->a[j].secondary_all = z[i];
                    a[i]->a[z[i]].secondary_all = -1;
                }
            }
            for (i = 0; i < 2; ++i) {
                int has_alt_hit = 0;
                if (n_pri[i] < a[i]->n) {
                    const AlnReg &p = a[i]->a[n_pri[i]];
                    if (!(p.score < opt.T || p.secondary >= 0 || !p.is_alt)) has_alt_hit = 1;
                }
                ro[i]->aln_off = ar.alloc((uint32_t)(1 + has_alt_hit) * (uint32_t)sizeof(AlnOut), err);
                ro[i]->n_aln = 0;
                if (*err == ERR_ARENA_OVERFLOW) return;
                AlnOut *out = reinterpret_cast<AlnOut *>(ar.base + ro[i]->aln_off);
                if (!(opt.flag & F_ALL)) xa_prepare(opt, rc[i], ws);
                AlnTmp h;
                aln_head(opt, mt, &a[i]->a[z[i]], h, err);
                h.mapq = q_se[i];
                h.flag |= 0x40 << i | extra_flag;
                aln_store(h, out[0]);
                push_task(tl, a[i]->a[z[i]], first + i, 0, ro[i]->aln_off, first + (i ^ 1), err); // also h[i], the mate record of the other end
                xa_emit(opt, rc[i], z[i], ws, ar, tl, ro[i]->aln_off, err);
                ro[i]->n_aln = 1;
                if (has_alt_hit) {
                    AlnTmp g;
                    aln_head(opt, mt, &a[i]->a[n_pri[i]], g, err);
                    g.flag |= 0x800 | 0x40 << i | extra_flag;
                    aln_store(g, out[1]);
                    const uint32_t off = ro[i]->aln_off + (uint32_t)sizeof(AlnOut);
                    push_task(tl, a[i]->a[n_pri[i]], first + i, 0, off, -1, err);
                    xa_emit(opt, rc[i], n_pri[i], ws, ar, tl, off, err);
                    ro[i]->n_aln = 2;
                }
            }
        }
    }
    if (paired) return;

    // no_pairing
    int h_rid[2];
    for (i = 0; i < 2; ++i) {
        int which = -1;
        if (a[i]->n) {
            if (a[i]->a[0].score >= opt.T) which = 0;
            else if (n_pri[i] < a[i]->n && a[i]->a[n_pri[i]].score >= opt.T) which = n_pri[i];
        }
        // h[i] is the first record emit_read() produces for this end (or unmapped); its rid is the region's
        h_rid[i] = which >= 0 && a[i]->a[which].rb >= 0 && a[i]->a[which].re >= 0 ? a[i]->a[which].rid : -1;
    }
    if (!(opt.flag & F_NOPAIRING) && h_rid[0] == h_rid[1] && h_rid[0] >= 0) {
        int64_t dist;
        int d = infer_dir(ix.l_pac, a[0]->a[0].rb, a[1]->a[0].rb, &dist);
        if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
    }
    emit_read(opt, mt, rc[0], 0x41 | extra_flag, first + 1, ws, ar, tl, reads, err);
    emit_read(opt, mt, rc[1], 0x81 | extra_flag, first, ws, ar, tl, reads, err);
}

} // namespace bsb
