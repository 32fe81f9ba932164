// bsb_extlane.h -- K5 as a per-read machine: one LANE owns one read, the warp meets at the DP row.
//
// Semantics: mem_chain2aln (bwamem.c:636-790) over the kept chains of a read, every extension = ksw_extend2
// (ksw.c:380-479) with this build's exhausted-row stop (bsb_ksw.h). The warp-per-read form (bsb_warp.cuh) spreads one DP row
// over the lanes: a max-scan, a REDUX and several ballots per 32 cells, with an average band of ~15 cells -- about 230 thread
// instructions per DP cell (profiles/r01_ncu_head.md). Here every lane runs the plain scalar recurrence of its OWN read
// (about 20 instructions per cell, no cross-lane traffic at all) and the control flow of mem_chain2aln is cut into a small
// state machine so that the 32 lanes of a warp, whatever stage their reads are in, execute the expensive part -- one DP row
// each -- together:
//
//     advance()   everything between two rows: next chain (window, seed order), next seed (containment tests), set-up of the
//                 left / right extension and its band retry, the finished region. Divergent, but short.
//     step()      up to C cells of the current row of the current extension (or of its initial row), with the row's closing
//                 logic when the chunk reaches the end of the band. All lanes in state ROW run it in the same loop.
//
// The (h,e) row of a lane lives in shared memory, lane-interleaved (word j * 32 + lane: conflict-free whatever j each lane is
// at), h and e packed in 16 bits each -- the launcher takes this kernel only when every score fits.
//
// No warp intrinsic is used in this file: the same machine is driven one read at a time by the CPU harness (tests/hostsim,
// HOSTSIM_EXT_LANES=1) against the reference's SAM, and 32 at a time by k_extend_lanes (bsb_cuda.cu).
#pragma once
#include "bsb_extend.h"

namespace bsb {

// Row storage of one lane: entry j of the row = word j * STRIDE (the kernel passes base + lane, STRIDE 32). One word per
// column: H (14 bits) | E (14 bits) << 14 | query base of that column << 28 -- the cell loop gets everything it needs from
// one shared-memory load; the launcher takes this kernel only when every score stays below 2^14.
constexpr int XL_BITS = 14;
constexpr uint32_t XL_MASK = (1u << XL_BITS) - 1u;
template <int STRIDE>
struct PackedRow {
    uint32_t *p;
    BSB_HD uint32_t *at(int j) const { return p + j * STRIDE; }
    static constexpr int stride = STRIDE;
};

// what advance() needs per query length, tabulated once per launch (double-precision divisions otherwise): cal_max_gap and
// the band limit of ksw_extend2 (min of max_ins, max_del) for the two end bonuses
struct ExtTables {
    const int *gap, *wl, *wr;   // [0, n)
    int n, amax;
};
BSB_HD int ext_band_limit(const Opt &opt, int qlen, int amax, int end_bonus)
{
    int max_ins = (int)((double)(qlen * amax + end_bonus - opt.o_ins) / opt.e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    int max_del = (int)((double)(qlen * amax + end_bonus - opt.o_del) / opt.e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    return max_ins < max_del ? max_ins : max_del;
}
BSB_HD int ext_amax(const Opt &opt)
{
    int mx = 0;
    for (int q = 0; q < 25; ++q) mx = mx > opt.mat[q] ? mx : opt.mat[q];
    return mx;
}
// fills one entry of the three tables (q < n); the kernel spreads the entries over the threads of a block
BSB_HD void ext_tables_fill(const Opt &opt, int amax, int q, int *gap, int *wl, int *wr)
{
    gap[q] = cal_max_gap(opt, q);
    wl[q] = ext_band_limit(opt, q, amax, opt.pen_clip5);
    wr[q] = ext_band_limit(opt, q, amax, opt.pen_clip3);
}

// DPX forms on the device (VIMNMX3, VIADDMNMX.RELU: one instruction each), plain C on the host
BSB_HD int bsb_max(int a, int b) { return a > b ? a : b; }
BSB_HD int bsb_min(int a, int b) { return a < b ? a : b; }
BSB_HD int bsb_max3(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
    return __vimax3_s32(a, b, c);
#else
    return bsb_max(bsb_max(a, b), c);
#endif
}
BSB_HD int bsb_addmax_relu(int a, int b, int c)   // max(a + b, c, 0)
{
#ifdef __CUDA_ARCH__
    return __viaddmax_s32_relu(a, b, c);
#else
    return bsb_max(bsb_max(a + b, c), 0);
#endif
}

template <class Row>
struct ExtLane {
    enum : int { IDLE = 0, NEXT_CHAIN, NEXT_SEED, EXT_SETUP, ROW, EXT_DONE, AFTER_LEFT, AFTER_RIGHT };
    int state;
    // read
    int r, len, nc, ci, err;
    uint32_t so;
    const uint8_t *seq;
    RegList av;
    // chain
    Chain c;
    const Seed *cs;
    uint64_t *srt;
    int64_t rmax0, rmax1;
    int k;
    // seed / region under construction
    Seed s;
    AlnReg a;
    int aw0, aw1, mo0, mo1, side, attempt, prev, sc0;
    // extension
    const uint8_t *qbase; int qdir;
    int64_t tstart; int tdir;
    int qlen, tlen, h0, w, zdrop, amax;
    int i, beg, end, max, max_i, max_j, max_ie, gscore, max_off;
    // row in progress: next column (-1: not opened), target base, H(i, j-1), F, row maximum key, first / last nonzero cell, potential
    int jc, rt, rh1, rf, rkey, rfirst, rlast, rphi;
    uint32_t n_cells;                            // DP cells this lane has filled (measurement: GCUPS of the launch)
    uint32_t tcache; int tc0, tcs;               // sixteen target bases (two bits each, first at the top) and how row i indexes them
    Row H; int row_cap;                          // words of this lane's row tile
    ExtTables T;

    BSB_HD int max_gap(const Opt &opt, int q) const { return (unsigned)q < (unsigned)T.n ? T.gap[q] : cal_max_gap(opt, q); }

    BSB_HD void begin_read(const BatchDev &B, int read)
    {
        r = read; err = 0;
        so = B.seed_off[r];
        const int ns = (int)(B.seed_off[r + 1] - so);
        if (ns == 0 || B.err[r]) { B.n_regs[r] = 0; state = IDLE; return; }
        len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
        seq = B.seq + B.seq_off[r];
        av.a = B.regs + so; av.n = 0; av.cap = ns;
        nc = B.n_chain[r]; ci = 0;
        srt = B.srt + so;
        state = NEXT_CHAIN;
    }

    BSB_HD void end_read(const BatchDev &B)
    {
        if (err) { B.err[r] = err; B.n_regs[r] = 0; }
        else B.n_regs[r] = av.n;              // regions before mem_sort_dedup_patch: the tail kernel finishes the read
        state = IDLE;
    }

    // Runs the control flow up to the next DP row (state ROW) or the end of the read (state IDLE).
    BSB_HD void advance(const Opt &opt, const IndexView &ix, const BatchDev &B, int max_q)
    {
        const int64_t l_pac = ix.l_pac;
        for (;;) {
            switch (state) {
            case NEXT_CHAIN: {
                if (err || ci >= nc) { end_read(B); return; }
                c = B.chains[so + ci];
                cs = B.cseeds + so + c.head;
                ++ci;
                if (c.n == 0) break;
                rmax0 = l_pac << 1; rmax1 = 0;
                for (int q = 0; q < c.n; ++q) {
                    const Seed &t = cs[q];
                    const int64_t b = t.rbeg - (t.qbeg + max_gap(opt, t.qbeg));
                    const int64_t e = t.rbeg + t.len + ((len - t.qbeg - t.len) + max_gap(opt, len - t.qbeg - t.len));
                    rmax0 = rmax0 < b ? rmax0 : b;
                    rmax1 = rmax1 > e ? rmax1 : e;
                }
                rmax0 = rmax0 > 0 ? rmax0 : 0;
                rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
                if (rmax0 < l_pac && l_pac < rmax1) {
                    if (cs[0].rbeg < l_pac) rmax1 = l_pac;
                    else rmax0 = l_pac;
                }
                fetch_window(ix, &rmax0, cs[0].rbeg, &rmax1);
                if (len > max_q) { err = ERR_SCRATCH_OVERFLOW; break; }
                for (int q = 0; q < c.n; ++q) srt[q] = (uint64_t)cs[q].score << 32 | (uint32_t)q;
                introsort((long)c.n, srt, LtU64());
                k = c.n - 1;
                state = NEXT_SEED;
                break;
            }
            case NEXT_SEED: {
                if (k < 0) { state = NEXT_CHAIN; break; }
                s = cs[(uint32_t)srt[k]];
                int q;
                for (q = 0; q < av.n; ++q) {       // is the seed inside a region found before?
                    const AlnReg &p = av.a[q];
                    int64_t rd;
                    int qd, w_, mg;
                    if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
                    if (s.len - p.seedlen0 > .1 * len) continue;
                    qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
                    mg = max_gap(opt, qd < rd ? qd : (int)rd);
                    w_ = mg < p.w ? mg : p.w;
                    if (qd - rd < w_ && rd - qd < w_) break;
                    qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
                    mg = max_gap(opt, qd < rd ? qd : (int)rd);
                    w_ = mg < p.w ? mg : p.w;
                    if (qd - rd < w_ && rd - qd < w_) break;
                }
                if (q < av.n) {                    // ... then extend it only if it overlaps a longer seed off-diagonal
                    for (q = k + 1; q < c.n; ++q) {
                        if (srt[q] == 0) continue;
                        const Seed &t = cs[(uint32_t)srt[q]];
                        if (t.len < s.len * .95) continue;
                        if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                        if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
                    }
                    if (q == c.n) { srt[k] = 0; --k; break; }
                }
                if (av.n >= av.cap) { err = ERR_SCRATCH_OVERFLOW; state = NEXT_CHAIN; break; }
                alnreg_clear(a);
                a.w = aw0 = aw1 = opt.w;
                a.score = a.truesc = -1;
                a.rid = c.rid;
                mo0 = mo1 = 0;
                if (s.qbeg) { side = 0; attempt = 0; state = EXT_SETUP; }
                else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; state = AFTER_LEFT; }
                break;
            }
            case EXT_SETUP: {
                int w_in, end_bonus;
                prev = a.score;
                if (side == 0) {
                    aw0 = w_in = opt.w << attempt;
                    qlen = s.qbeg; qbase = seq + (s.qbeg - 1); qdir = -1;
                    tstart = s.rbeg - 1; tdir = -1; tlen = (int)(s.rbeg - rmax0);
                    end_bonus = opt.pen_clip5; h0 = s.len * opt.a;
                } else {
                    aw1 = w_in = opt.w << attempt;
                    const int qe = s.qbeg + s.len;
                    const int64_t re = s.rbeg + s.len - rmax0;
                    qlen = len - qe; qbase = seq + qe; qdir = 1;
                    tstart = rmax0 + re; tdir = 1; tlen = (int)(rmax1 - rmax0 - re);
                    end_bonus = opt.pen_clip3; h0 = sc0;
                }
                amax = T.amax;
                const int lim = (unsigned)qlen < (unsigned)T.n ? (side == 0 ? T.wl[qlen] : T.wr[qlen]) : ext_band_limit(opt, qlen, amax, end_bonus);
                w = w_in < lim ? w_in : lim;
                zdrop = opt.zdrop;
                max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
                beg = 0; end = qlen;
                if (qlen + 2 > row_cap) { err = ERR_ROW_TILE; state = NEXT_CHAIN; break; }
                i = -1; jc = 0;                    // init_step() fills the initial row
                state = ROW;
                return;
            }
            case EXT_DONE: {
                a.score = max;
                const int aw = side == 0 ? aw0 : aw1;
                if (side == 0) mo0 = max_off; else mo1 = max_off;
                if (attempt == 0 && !(a.score == prev || max_off < (aw >> 1) + (aw >> 2))) { attempt = 1; state = EXT_SETUP; break; }
                const int qle = max_j + 1, tle = max_i + 1, gtle = max_ie + 1;
                if (side == 0) {
                    if (gscore <= 0 || gscore <= a.score - opt.pen_clip5) { a.qb = s.qbeg - qle; a.rb = s.rbeg - tle; a.truesc = a.score; }
                    else { a.qb = 0; a.rb = s.rbeg - gtle; a.truesc = gscore; }
                    state = AFTER_LEFT;
                } else {
                    const int qe = s.qbeg + s.len;
                    const int64_t re = s.rbeg + s.len - rmax0;
                    if (gscore <= 0 || gscore <= a.score - opt.pen_clip3) { a.qe = qe + qle; a.re = rmax0 + re + tle; a.truesc += a.score - sc0; }
                    else { a.qe = len; a.re = rmax0 + re + gtle; a.truesc += gscore - sc0; }
                    state = AFTER_RIGHT;
                }
                break;
            }
            case AFTER_LEFT: {
                if (s.qbeg + s.len != len) { sc0 = a.score; side = 1; attempt = 0; state = EXT_SETUP; }
                else { a.qe = len; a.re = s.rbeg + s.len; state = AFTER_RIGHT; }
                break;
            }
            case AFTER_RIGHT: {
                a.seedcov = 0;
                for (int q = 0; q < c.n; ++q) {
                    const Seed &t = cs[q];
                    if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
                }
                a.w = aw0 > aw1 ? aw0 : aw1;
                a.seedlen0 = s.len;
                a.frac_rep = c.frac_rep;
                av.a[av.n++] = a;
                --k;
                state = NEXT_SEED;
                break;
            }
            default:
                return;                            // IDLE, ROW
            }
        }
    }

    // The sixteen target bases of rows [i, i + 16) in one word. The rows walk the doubled coordinate space in direction tdir;
    // on the forward strand that is the 2-bit pac itself, on the reverse strand the complement of the mirrored position
    // (ref_base, bsb_index.h), so the sixteen bases are always a run of consecutive pac positions, ascending or descending.
    BSB_HD void fetch_target(const IndexView &ix)
    {
        const int n = tlen - i < 16 ? tlen - i : 16;
        const int64_t l2 = ix.l_pac << 1;
        const int64_t p0 = tstart + (int64_t)i * tdir, pl = p0 + (int64_t)(n - 1) * tdir;
        const bool rev = p0 >= ix.l_pac;
        const int64_t f0 = rev ? l2 - 1 - p0 : p0, fl = rev ? l2 - 1 - pl : pl;
        const int64_t fs = f0 < fl ? f0 : fl;
        const uint8_t *b = ix.pac + (fs >> 2);
        const int need = ((int)(fs & 3) + n + 3) >> 2;   // bytes that hold the n bases: 1..5
        uint64_t x = (uint64_t)b[0] << 56;
        if (need > 1) x |= (uint64_t)b[1] << 48;
        if (need > 2) x |= (uint64_t)b[2] << 40;
        if (need > 3) x |= (uint64_t)b[3] << 32;
        if (need > 4) x |= (uint64_t)b[4] << 24;
        x <<= 2 * (int)(fs & 3);
        uint32_t t32 = (uint32_t)(x >> 32);           // base fs + k at bits 31-2k, 30-2k
        if (rev) t32 = ~t32;
        tcache = t32;
        if (f0 <= fl) { tc0 = 0; tcs = 1; } else { tc0 = n - 1; tcs = -1; }
    }

    // One chunk of the initial row (state ROW, i == -1): H[0] = h0, H[j] = h0 - oe_ins - (j - 1) e_ins while positive, E = 0,
    // and the query base of column j into the same word. The kernel runs this for all lanes that have just set up an
    // extension, right after the control phase, so that they do it together.
    BSB_HD void init_step(const Opt &opt, int C)
    {
        const int oe_ins = opt.o_ins + opt.e_ins, e_ins = opt.e_ins;
        int j = jc;
        const int stop_at = j + C < qlen + 1 ? j + C : qlen + 1;
        uint32_t *hp = H.at(j);
        const uint8_t *qp = qbase + (long)j * qdir;
        int v = j == 0 ? h0 : h0 - oe_ins - (j - 1) * e_ins;
        for (; j < stop_at; ++j, hp += Row::stride, qp += qdir) {
            v = v > 0 ? v : 0;
            const uint32_t q = j < qlen ? *qp : 4;
            *hp = (uint32_t)v | q << (2 * XL_BITS);
            v = j == 0 ? h0 - oe_ins : v - e_ins;
        }
        jc = j;
        if (j > qlen) { i = 0; jc = -1; if (tlen <= 0) state = EXT_DONE; }
    }
    BSB_HD bool in_init() const { return state == ROW && i < 0; }

    // Up to C cells of the current row; a lane that reaches the end of its row closes it (row maximum, to-end score, z-drop,
    // next band) and starts the next one at the following call. Rows are cut into chunks because the band widths of the 32
    // reads of a warp differ by an order of magnitude: with whole rows the warp waits for its widest. Leaves state ROW when
    // the extension ends.
    BSB_HD void step(const Opt &opt, const IndexView &ix, int C)
    {
        const int oe_del = opt.o_del + opt.e_del, oe_ins = opt.o_ins + opt.e_ins, e_del = opt.e_del, e_ins = opt.e_ins;
        if (i < 0) { init_step(opt, 4 * C); return; }
        if (jc < 0) {                              // open row i
            if ((i & 15) == 0) fetch_target(ix);
            rt = (int)(tcache >> (30 - 2 * (tc0 + tcs * (i & 15)))) & 3;
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
            if (beg == 0) { rh1 = h0 - (opt.o_del + e_del * (i + 1)); if (rh1 < 0) rh1 = 0; }
            else rh1 = 0;
            rf = 0; rkey = -1; rfirst = 0x7fff; rlast = -1; rphi = 0;
            jc = beg;
        }
        {
            // bwa_fill_scmat: match a, mismatch -b, anything against an ambiguous base -1 (the launcher checks opt.mat has this
            // form; the target comes from the 2-bit pac and is never ambiguous)
            const int t = rt, s_match = opt.a, s_mis = -opt.b;
            const bool at_qend = end == qlen;
            int j = jc, h1 = rh1, f = rf, key = rkey, first_nz = rfirst, last_nz = rlast, phi = rphi;
            const int stop_at = j + C < end ? j + C : end;
            n_cells += (uint32_t)(stop_at - j);
            uint32_t *hp = H.at(j);
            int A = at_qend ? amax * (qlen - j) : 0;
            const int dA = at_qend ? amax : 0;         // away from the query end the potential is not needed: phi stays 0
#define BSB_XL_CELL(K)                                                                                                   \
            {                                                                                                           \
                const uint32_t p = hp[(K) * Row::stride];                                                               \
                const int Mo = (int)(p & XL_MASK), eo = (int)(p >> XL_BITS & XL_MASK), q = (int)(p >> (2 * XL_BITS));    \
                const int sc = q == t ? s_match : (q > 3 ? -1 : s_mis);                                                 \
                const int stored_h = h1;                                                                                \
                /* a negative M acts like 0 everywhere below (h takes the maximum with e, f >= 0; gap sources are clamped) */ \
                int M = Mo + sc; M = M > 0 ? M : 0; M = Mo ? M : 0;                                                     \
                const int h = bsb_max3(M, eo, f);                                                                       \
                h1 = h;                                                                                                 \
                key = bsb_max(key, h * 4096 + (j + (K)));          /* row maximum; among equal cells the last column */ \
                const int e = bsb_addmax_relu(eo, -e_del, M - oe_del);                                                  \
                f = bsb_addmax_relu(f, -e_ins, M - oe_ins);                                                             \
                hp[(K) * Row::stride] = (p & ~((1u << 2 * XL_BITS) - 1u)) | (uint32_t)stored_h | (uint32_t)e << XL_BITS; \
                const int v = bsb_max(stored_h, e);   /* the stored cell: nonzero? and what a later row can make of it */ \
                const bool nz = v != 0;                                                                                 \
                last_nz = nz ? j + (K) : last_nz;                                                                       \
                first_nz = bsb_min(first_nz, nz ? j + (K) : 0x7fff);                                                    \
                phi = bsb_max(phi, nz ? v + A - (K) * dA : 0);   /* >= the potential of bsb_ksw.h (E gets one column more) */ \
            }
            for (; j + 4 <= stop_at; j += 4, hp += 4 * Row::stride, A -= 4 * dA) {
                BSB_XL_CELL(0) BSB_XL_CELL(1) BSB_XL_CELL(2) BSB_XL_CELL(3)
            }
            for (; j < stop_at; ++j, hp += Row::stride, A -= dA) BSB_XL_CELL(0)
#undef BSB_XL_CELL
            jc = j; rh1 = h1; rf = f; rkey = key; rfirst = first_nz; rlast = last_nz; rphi = phi;
            if (j < end) return;
        }
        // close row i
        const int h1 = rh1;
        const int m = rkey < 0 ? 0 : rkey >> 12, mj = rkey < 0 ? -1 : rkey & 0xfff;
        const int first_nz = rfirst == 0x7fff ? -1 : rfirst, last_nz = rlast;
        const bool at_qend = end == qlen;
        jc = -1;
        {   // H[end] = h1, E[end] = 0 (the query base of that column stays)
            uint32_t *hp = H.at(end);
            *hp = (*hp & ~((1u << 2 * XL_BITS) - 1u)) | (uint32_t)h1;
        }
        if ((beg < end ? end : beg) == qlen) {     // `j == qlen` after the reference's loop (j stays at beg when the band is empty)
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        bool stop = m == 0;
        if (!stop) {
            if (m > max) {
                max = m; max_i = i; max_j = mj;
                const int off = iabs(mj - i);
                max_off = max_off > off ? max_off : off;
            } else if (zdrop > 0) {
                if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) stop = true; }
                else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) stop = true; }
            }
        }
        if (!stop && at_qend && ext_rows_exhausted(rphi, max, gscore)) stop = true;
        if (!stop) {
            // next band: drop leading / trailing cells whose h and e are both zero
            const int nb = first_nz >= 0 ? first_nz : end;
            int jl;
            if (h1 != 0) jl = end;
            else if (last_nz >= 0 && last_nz >= nb) jl = last_nz;
            else jl = nb - 1;
            beg = nb;
            end = jl + 2 < qlen ? jl + 2 : qlen;
            if (++i >= tlen) stop = true;
        }
        if (stop) state = EXT_DONE;
    }
};

// The tail of mem_align1_core for one read, after the machine: mem_sort_dedup_patch (bwamem.c:441-493) and the ALT marks.
BSB_HD void extend_tail(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, DpScratch &dp)
{
    int n = B.n_regs[r];
    if (n <= 0 || B.err[r]) return;
    AlnReg *a = B.regs + B.seed_off[r];
    int err = 0;
    n = sort_dedup_patch(opt, ix, B.seq + B.seq_off[r], n, a, dp, &err);
    for (int q = 0; q < n; ++q)
        if (a[q].rid >= 0 && ix.anns[a[q].rid].is_alt) a[q].is_alt = 1;
    if (err) { B.err[r] = err; B.n_regs[r] = 0; }
    else B.n_regs[r] = n;
}

} // namespace bsb
