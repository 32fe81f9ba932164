// bsb_extend.h -- chain -> alignment regions for one read (north_star stage 5).
//
//   global_core()        <- bwa_gen_cigar2 part 1       (bwa.c:199-249)  score (and CIGAR) between fixed end points
//   (mem_chain2aln, bwamem.c:636-790: the lane machine of bsb_extlane.h and chain_to_regions_warp of bsb_warp.cuh; a plain
//    scalar form for the CPU harness lives in tests/hostsim/scalar_stages.h)
//   patch_regions()      <- mem_patch_reg               (bwamem.c:410-439)
//   sort_dedup_patch()   <- mem_sort_dedup_patch        (bwamem.c:441-493)
// Attribution: restates BWA-MEM's mem_sort_dedup_patch, mem_patch_reg (bwamem.c) and bwa_gen_cigar2 (bwa.c); GPLv3, Heng Li.
// See NOTICE.md.
#pragma once
#include "bsb_ksw.h"
#include "bsb_chain.h"

namespace bsb {

BSB_HD int cal_max_gap(const Opt &opt, int qlen)
{
    int l_del = (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
    int l_ins = (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < opt.w << 1 ? l : opt.w << 1;
}

struct DpScratch {        // per-thread DP scratch in HBM
    int32_t *eh;          // 2*(max_q+1)
    uint8_t *z;           // traceback, z_cap bytes
    long z_cap;
    int max_q;
};

// Global alignment of query[0,l_query) against reference [rb,re) (doubled coordinates).
// Returns false when the reference rejects the request (bwa.c:212, 215). When `cig` is null only
// the score is produced (this is what region patching needs).
BSB_HD bool global_core(const Opt &opt, const IndexView &ix, int w_, int l_query, const uint8_t *query,
                        int64_t rb, int64_t re, int *score, CigarBuf *cig, DpScratch &dp, int *err)
{
    const int64_t l_pac = ix.l_pac;
    if (cig) cig->n = 0;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
    // bns_get_seq clamps to [0, 2*l_pac); a clamped fetch has the wrong length and is rejected
    if (re > (l_pac << 1) || rb < 0) return false;
    int64_t rlen = re - rb;
    QrySeq q; RefSeq t;
    if (rb >= l_pac) { // reverse both so that gaps are left-aligned on the forward strand
        q.base = query + (l_query - 1); q.dir = -1;
        t.pac = ix.pac; t.l_pac = l_pac; t.start = re - 1; t.dir = -1;
    } else {
        q.base = query; q.dir = 1;
        t.pac = ix.pac; t.l_pac = l_pac; t.start = rb; t.dir = 1;
    }
    if (l_query == rlen && w_ == 0) {
        if (cig) cig->push(0, l_query);
        int sc = 0;
        for (int i = 0; i < l_query; ++i) sc += opt.mat[t(i) * 5 + q(i)];
        *score = sc;
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_ins) / opt.e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_del) / opt.e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        w = (max_gap + iabs((int)rlen - l_query) + 1) >> 1;
        w = w < w_ ? w : w_;
        min_w = iabs((int)rlen - l_query) + 3;
        w = w > min_w ? w : min_w;
        if (l_query > dp.max_q) { *err = ERR_SCRATCH_OVERFLOW; return false; }
        if (cig) {
            long n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
            if (n_col * rlen > dp.z_cap) { *err = ERR_SCRATCH_OVERFLOW; return false; }
        }
        *score = nw_global(l_query, q, (int)rlen, t, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, dp.eh, dp.z, cig, err);
    }
    return true;
}

BSB_HD void alnreg_clear(AlnReg &a)
{
    a.rb = a.re = 0; a.qb = a.qe = 0; a.rid = 0; a.score = a.truesc = a.sub = a.alt_sc = a.csub = a.sub_n = 0;
    a.w = a.seedcov = a.secondary = a.secondary_all = a.seedlen0 = a.n_comp = a.is_alt = 0;
    a.frac_rep = 0; a.hash = 0;
}

struct RegList { AlnReg *a; int n, cap; };

// Extends the seeds of one chain (longest first) and appends the regions to av.
// mem_flt_chained_seeds' threshold (bwamem.c:604-606): the minimum seed score, or -1 when the filter does not run for this
// read length (5.5 ln(l) > 0.05 l, i.e. every read shorter than about 720 bp unless -W is given). Float/double mix as in C:
// MEM_HSP_COEF * int and MEM_SEEDSW_COEF * int are float products; log_tab[i] = glibc log(i).
BSB_HD int seed_sw_min_score(const Opt &opt, int l_query, const double *log_tab, int n_log, int *err)
{
    double min_l;
    if (opt.min_chain_weight) min_l = (double)(1.1f * (float)opt.min_chain_weight);
    else {
        if (l_query >= n_log) { *err = ERR_SCRATCH_OVERFLOW; return -1; }
        min_l = (double)5.5f * log_tab[l_query];
    }
    if (min_l > (double)(0.05f * (float)l_query)) return -1;
    return (int)(opt.a * min_l + .499);
}

// mem_seed_sw (bwamem.c:575-600): local alignment score of the seed with 50 bp of flank on either side, -1 when the seed
// (or its window) is 200 bp or longer. ksw_align2 without KSW_XBYTE runs the 16-bit kernel; only the score is used.
BSB_HD int seed_sw(const Opt &opt, const IndexView &ix, int l_query, const uint8_t *query, const Seed &s, SwScratch &ws, int *err)
{
    const int64_t l_pac = ix.l_pac;
    if (s.len >= 200) return -1;
    int qb = s.qbeg, qe = s.qbeg + s.len;
    int64_t rb = s.rbeg, re = s.rbeg + s.len;
    const int64_t mid = (rb + re) >> 1;
    qb -= 50; qb = qb > 0 ? qb : 0;
    qe += 50; qe = qe < l_query ? qe : l_query;
    rb -= 50; rb = rb > 0 ? rb : 0;
    re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
    if (rb < l_pac && l_pac < re) {
        if (mid < l_pac) re = l_pac;
        else rb = l_pac;
    }
    if (qe - qb >= 200 || re - rb >= 200) return -1;
    fetch_window(ix, &rb, mid, &re);
    QrySeq q = {query + qb, 1};
    RefSeq t = {ix.pac, l_pac, rb, 1};
    SwResult x = sw_striped(2, qe - qb, q, (int)(re - rb), t, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, SW_XSTART, ws, err);
    return x.score;
}

// mem_flt_chained_seeds (bwamem.c:602-619) over the compacted chains of one read: seeds whose flanked local score is below
// the threshold leave their chain, the others carry that score into the extension order (chain_to_regions sorts by it).
BSB_HD void filter_chained_seeds(const Opt &opt, const IndexView &ix, int l_query, const uint8_t *query, int n_chn, Chain *chains,
                                 Seed *cseeds, int min_hsp, SwScratch &ws, int *err)
{
    for (int i = 0; i < n_chn; ++i) {
        Chain &c = chains[i];
        Seed *cs = cseeds + c.head;
        int k = 0;
        for (int j = 0; j < c.n; ++j) {
            Seed s = cs[j];
            int sc = seed_sw(opt, ix, l_query, query, s, ws, err);
            if (*err) return;
            if (sc < 0 || sc >= min_hsp) {
                s.score = sc < 0 ? s.len * opt.a : sc;
                cs[k++] = s;
            }
        }
        c.n = k;
    }
}


// Can regions a (left) and b (right) be one alignment? Returns the merged score or 0.
BSB_HD int patch_regions(const Opt &opt, const IndexView &ix, const uint8_t *query, const AlnReg &a, const AlnReg &b, int *_w, DpScratch &dp, int *err)
{
    int w, score = 0, q_s, r_s;
    double r;
    if (query == nullptr) return 0;
    if (a.rb < ix.l_pac && b.rb >= ix.l_pac) return 0;
    if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return 0;
    w = (int)((a.re - b.rb) - (a.qe - b.qb));
    w = w > 0 ? w : -w;
    r = (double)(a.re - b.rb) / (b.re - a.rb) - (double)(a.qe - b.qb) / (b.qe - a.qb);
    r = r > 0. ? r : -r;
    if (a.re < b.rb || a.qe < b.qb) {
        if (w > opt.w << 1 || r >= 0.05f) return 0;
    } else if (w > opt.w << 2 || r >= 0.05f * 2) return 0;
    w += a.w + b.w;
    w = w < opt.w << 2 ? w : opt.w << 2;
    if (!global_core(opt, ix, w, b.qe - a.qb, query + a.qb, a.rb, b.re, &score, nullptr, dp, err)) return 0;
    q_s = (int)((double)(b.qe - a.qb) / ((b.qe - b.qb) + (a.qe - a.qb)) * (b.score + a.score) + .499);
    r_s = (int)((double)(b.re - a.rb) / ((b.re - b.rb) + (a.re - a.rb)) * (b.score + a.score) + .499);
    if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
    *_w = w;
    return score;
}

struct LtRegRe { BSB_HD bool operator()(const AlnReg &a, const AlnReg &b) const { return a.re < b.re; } };
struct LtRegScore {
    BSB_HD bool operator()(const AlnReg &a, const AlnReg &b) const
    { return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb))); }
};

// query == nullptr disables patching (mate-rescue call site, bwamem_pair.c:175)
BSB_HD int sort_dedup_patch(const Opt &opt, const IndexView &ix, const uint8_t *query, int n, AlnReg *a, DpScratch &dp, int *err)
{
    int m, i, j;
    if (n <= 1) return n;
    introsort((long)n, a, LtRegRe());
    for (i = 0; i < n; ++i) a[i].n_comp = 1;
    for (i = 1; i < n; ++i) {
        AlnReg &p = a[i];
        if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt.max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p.rid == a[j].rid && p.rb < a[j].re + opt.max_chain_gap; --j) {
            AlnReg &q = a[j];
            int64_t orr, oq, mr, mq;
            int score, w;
            if (q.qe == q.qb) continue;
            orr = q.re - p.rb;
            oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
            mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
            mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
            if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
                if (p.score < q.score) { p.qe = p.qb; break; }
                else q.qe = q.qb;
            } else if (q.rb < p.rb && (score = patch_regions(opt, ix, query, q, p, &w, dp, err)) > 0) {
                p.n_comp += q.n_comp + 1;
                p.seedcov = p.seedcov > q.seedcov ? p.seedcov : q.seedcov;
                p.sub = p.sub > q.sub ? p.sub : q.sub;
                p.csub = p.csub > q.csub ? p.csub : q.csub;
                p.qb = q.qb; p.rb = q.rb;
                p.truesc = p.score = score;
                p.w = w;
                q.qb = q.qe;
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i)
        if (a[i].qe > a[i].qb) {
            if (m != i) a[m++] = a[i];
            else ++m;
        }
    n = m;
    introsort((long)n, a, LtRegScore());
    for (i = 1; i < n; ++i)
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) {
            if (m != i) a[m++] = a[i];
            else ++m;
        }
    return m;
}

} // namespace bsb
