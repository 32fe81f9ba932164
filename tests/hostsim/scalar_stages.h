// scalar_stages.h -- test infrastructure: plain one-read-at-a-time forms of the stages, used only by the CPU harness
// (tests/hostsim) to pin the kernel bodies against the reference's SAM in a container without a GPU. Nothing here is
// compiled into libbsbolt_b200.so. The product forms are k_seed3 (bsb_seed3.h), ExtLane (bsb_extlane.h) and the warp kernels of
// bsb_warp.cuh.
//
// The algorithms restated here are BWA-MEM's (Heng Li; bwa 0.7.17 as forked by BSBolt, GPLv3, bsbolt/External/BWA/):
//   collect_intv / smem_at / seed_forward <- mem_collect_intv, bwt_smem1a, bwt_seed_strategy1 (bwamem.c:118-166, bwt.c:285-383)
//   sw_extend                            <- ksw_extend2     (ksw.c:380-479)
//   chain_to_regions                     <- mem_chain2aln   (bwamem.c:636-790)
#pragma once
#include "../../bsbolt_b200/csrc/bsb_stages.h"
#include "bsb_smem.h"
#include "bsb_smem_sm.h"

namespace bsb {

struct SeedScratch { Intv *mem1, *t0, *t1; };

// K2: SMEM seeding for read r. use_sm selects the converged state-machine form (all lanes of a warp call
// together, `active` false for lanes without a read); both forms produce the same interval list.
BSB_HD void stage_seed(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, const SeedScratch &sc, bool use_sm = false, bool active = true, bool use_v3 = false)
{
    const int len = active ? (int)(B.seq_off[r + 1] - B.seq_off[r]) : 0;
    const uint8_t *seq = active ? B.seq + B.seq_off[r] : nullptr;
    IntvList mem = {active ? B.intv + (size_t)r * B.intv_cap : nullptr, 0, B.intv_cap};
    IntvList mem1 = {sc.mem1, 0, B.intv_cap}, t0 = {sc.t0, 0, B.intv_cap}, t1 = {sc.t1, 0, B.intv_cap};
    int err = 0;
    if (active) { B.n_intv[r] = 0; B.l_rep[r] = 0; B.n_seed[r] = 0; }
    const bool work = active && len >= opt.min_seed_len;
    if (use_v3) {   // the product's seeding form (k_seed3), driven sequentially: list storage borrowed from the scratch lists
        if (work) {
            ListPlain L = {(uint64_t *)sc.t0, (uint64_t *)sc.t1, (int *)sc.mem1, B.intv_cap};
            BasesBytes q = {seq};
            mem.n = ix.occ32 ? collect_intv_v3<uint32_t>(opt, ix, len, q, L, mem.a, mem.cap, &err)
                             : collect_intv_v3<uint64_t>(opt, ix, len, q, L, mem.a, mem.cap, &err);
        }
    } else if (use_sm) collect_intv_sm(opt, ix, len, seq, mem, mem1, t0, t1, &err, work);
    else if (work) collect_intv(opt, ix, len, seq, mem, mem1, t0, t1, &err);
    if (!work) return;
    seed_finish(opt, B, r, mem.a, mem.n, err);
}


// eh: scratch of 2*(qlen+1) ints
template <class Q, class T>
BSB_HD ExtResult sw_extend(int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                           int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                           int32_t *eh)
{
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
    int32_t *H = eh, *E = eh + (qlen + 1);
    for (j = 0; j <= qlen; ++j) H[j] = E[j] = 0;
    H[0] = h0; H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;
    for (i = 0, max = 0; i < 25; ++i) max = max > mat[i] ? max : mat[i];
    max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    const int amax = max;
    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
    beg = 0; end = qlen;
    for (i = 0; i < tlen; ++i) {
        int t, f = 0, h1, m = 0, mj = -1;
        const int8_t *row = mat + target(i) * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) {
            h1 = h0 - (o_del + e_del * (i + 1));
            if (h1 < 0) h1 = 0;
        } else h1 = 0;
        for (j = beg; j < end; ++j) {
            int h, M = H[j], e = E[j];
            H[j] = h1;
            M = M ? M + row[query(j)] : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            mj = m > h ? mj : j;
            m = m > h ? m : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            E[j] = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
        }
        H[end] = h1; E[end] = 0;
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            max_off = max_off > iabs(mj - i) ? max_off : iabs(mj - i);
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
            }
        }
        if (end == qlen) {
            int phi = 0;
            for (j = beg; j < end; ++j) {
                t = H[j] > 0 ? H[j] + amax * (qlen - j) : 0; phi = phi > t ? phi : t;
                t = E[j] > 0 ? E[j] + amax * (qlen - 1 - j) : 0; phi = phi > t ? phi : t;
            }
            if (ext_rows_exhausted(phi, max, gscore)) break;
        }
        for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    (void)k;
    ExtResult r;
    r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
    return r;
}


// cs: the chain's seeds (contiguous); srt: scratch u64[c.n]
BSB_HD void chain_to_regions(const Opt &opt, const IndexView &ix, int l_query, const uint8_t *query,
                             const Chain &c, const Seed *cs, uint64_t *srt, RegList &av, DpScratch &dp, int *err)
{
    int i, k, max_off[2], aw[2];
    const int64_t l_pac = ix.l_pac;
    int64_t rmax[2], tmp, max = 0;
    if (c.n == 0) return;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (i = 0; i < c.n; ++i) {
        const Seed &t = cs[i];
        int64_t b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
        int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
        if (t.len > max) max = t.len;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
    if (rmax[0] < l_pac && l_pac < rmax[1]) {
        if (cs[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    fetch_window(ix, &rmax[0], cs[0].rbeg, &rmax[1]);
    if (l_query > dp.max_q) { *err = ERR_SCRATCH_OVERFLOW; return; }

    for (i = 0; i < c.n; ++i) srt[i] = (uint64_t)cs[i].score << 32 | (uint32_t)i;
    introsort((long)c.n, srt, LtU64());

    for (k = c.n - 1; k >= 0; --k) {
        const Seed &s = cs[(uint32_t)srt[k]];
        for (i = 0; i < av.n; ++i) { // already covered by an earlier extension?
            const AlnReg &p = av.a[i];
            int64_t rd;
            int qd, w, max_gap;
            if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
            if (s.len - p.seedlen0 > .1 * l_query) continue;
            qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p.w ? max_gap : p.w;
            if (qd - rd < w && rd - qd < w) break;
            qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p.w ? max_gap : p.w;
            if (qd - rd < w && rd - qd < w) break;
        }
        if (i < av.n) {
            for (i = k + 1; i < c.n; ++i) { // an overlapping, non-colinear longer seed forces an extension
                if (srt[i] == 0) continue;
                const Seed &t = cs[(uint32_t)srt[i]];
                if (t.len < s.len * .95) continue;
                if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
            }
            if (i == c.n) { srt[k] = 0; continue; }
        }
        if (av.n >= av.cap) { *err = ERR_SCRATCH_OVERFLOW; return; }
        AlnReg &a = av.a[av.n++];
        alnreg_clear(a);
        a.w = aw[0] = aw[1] = opt.w;
        a.score = a.truesc = -1;
        a.rid = c.rid;

        if (s.qbeg) { // left extension over the reversed prefix
            QrySeq qs = {query + (s.qbeg - 1), -1};
            RefSeq rs = {ix.pac, l_pac, s.rbeg - 1, -1};
            tmp = s.rbeg - rmax[0];
            ExtResult x = {0, 0, 0, 0, 0, 0};
            for (i = 0; i < 2; ++i) {
                int prev = a.score;
                aw[0] = opt.w << i;
                x = sw_extend(s.qbeg, qs, (int)tmp, rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[0], opt.pen_clip5, opt.zdrop, s.len * opt.a, dp.eh);
                a.score = x.score; max_off[0] = x.max_off;
                if (a.score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
            }
            if (x.gscore <= 0 || x.gscore <= a.score - opt.pen_clip5) {
                a.qb = s.qbeg - x.qle; a.rb = s.rbeg - x.tle;
                a.truesc = a.score;
            } else {
                a.qb = 0; a.rb = s.rbeg - x.gtle;
                a.truesc = x.gscore;
            }
        } else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }

        if (s.qbeg + s.len != l_query) { // right extension
            int qe = s.qbeg + s.len, sc0 = a.score;
            int64_t re = s.rbeg + s.len - rmax[0];
            QrySeq qs = {query + qe, 1};
            RefSeq rs = {ix.pac, l_pac, rmax[0] + re, 1};
            ExtResult x = {0, 0, 0, 0, 0, 0};
            for (i = 0; i < 2; ++i) {
                int prev = a.score;
                aw[1] = opt.w << i;
                x = sw_extend(l_query - qe, qs, (int)(rmax[1] - rmax[0] - re), rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[1], opt.pen_clip3, opt.zdrop, sc0, dp.eh);
                a.score = x.score; max_off[1] = x.max_off;
                if (a.score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
            }
            if (x.gscore <= 0 || x.gscore <= a.score - opt.pen_clip3) {
                a.qe = qe + x.qle; a.re = rmax[0] + re + x.tle;
                a.truesc += a.score - sc0;
            } else {
                a.qe = l_query; a.re = rmax[0] + re + x.gtle;
                a.truesc += x.gscore - sc0;
            }
        } else { a.qe = l_query; a.re = s.rbeg + s.len; }

        a.seedcov = 0;
        for (i = 0; i < c.n; ++i) {
            const Seed &t = cs[i];
            if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
        }
        a.w = aw[0] > aw[1] ? aw[0] : aw[1];
        a.seedlen0 = s.len;
        a.frac_rep = c.frac_rep;
    }
}

// K5: banded extension of every kept chain of read r + region de-duplication
BSB_HD void stage_extend(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, DpScratch &dp)
{
    const uint32_t so = B.seed_off[r];
    const int ns = (int)(B.seed_off[r + 1] - so);
    B.n_regs[r] = 0;
    if (ns == 0 || B.err[r]) return;
    const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    const uint8_t *seq = B.seq + B.seq_off[r];
    RegList av = {B.regs + so, 0, ns};
    int err = 0;
    const int nc = B.n_chain[r];
    for (int i = 0; i < nc; ++i) {
        const Chain &c = B.chains[so + i];
        chain_to_regions(opt, ix, len, seq, c, B.cseeds + so + c.head, B.srt + so, av, dp, &err);
        if (err) { B.err[r] = err; return; }
    }
    av.n = sort_dedup_patch(opt, ix, seq, av.n, av.a, dp, &err);
    for (int i = 0; i < av.n; ++i)
        if (av.a[i].rid >= 0 && ix.anns[av.a[i].rid].is_alt) av.a[i].is_alt = 1;
    if (err) { B.err[r] = err; return; }
    B.n_regs[r] = av.n;
}


} // namespace bsb
