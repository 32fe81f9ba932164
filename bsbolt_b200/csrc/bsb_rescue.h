// bsb_rescue.h -- the Smith-Waterman of mate rescue (mem_matesw -> ksw_align2 -> ksw_u8, ksw.c:115-339), one LANE per job.
//
// The jobs of a batch (RescueJob, bsb_final.h) are independent of each other, all about the same shape (the mate against
// an insert-size window), and there are hundreds of thousands of them in a library with unpaired or mis-converted mates:
// exactly the case for the plain scalar recurrence per lane, like the extension kernel (bsb_extlane.h). The warp-per-pair
// form (sw_striped_warp, bsb_warp.cuh) spends ~150 thread instructions per cell on its two max-scans.
//
// Semantics are those of sw_striped() (bsb_ksw.h), the cell-exact linear restatement of the striped SSE2 kernel, for the
// 8-bit kernel (reads below 250 / a bases): the first sweep restarts F at the head of every stripe of slen cells, E and the
// row maximum come from that sweep, and the lazy-F repair -- F carried across the stripe boundaries, H only -- is fused
// into the same pass as a second running value (it only needs the first-sweep h of the cell and its own past). Zero-scoring
// pad cells, saturation at 255 - shift, the sub-optimal list b[] and the reverse pass of ksw_align2 are all reproduced.
//
// One 32-bit word per cell in shared memory, lane-interleaved: three H rows (bytes 0, 1, 3) and E (byte 2). Two of the H
// rows are the rows in flight; the third keeps the row of the best score -- when a row becomes the new best its byte simply
// changes role (it is only read from then on: as the previous row of the next one and as the kept row), nothing is copied.
// The query codes follow the cells, eight per word.
//
// No warp intrinsic in this file: tests/hostsim runs the same machine (HOSTSIM_RESCUE_JOBS=1) against the reference's SAM.
#pragma once
#include "bsb_final.h"
#include "bsb_extlane.h"

namespace bsb {

template <class Row>
struct SwLane {
    enum : int { IDLE = 0, INIT, ROWS, PASS_END };
    int state;
    // job
    const uint8_t *pac; int64_t l_pac, rb; const uint8_t *ms; int l_ms, is_rev, tlen;
    int min_score;                 // opt.min_seed_len * opt.a: the XSUBO threshold of the forward pass
    // current pass
    int pass, qlen, slen, L, minsc, endsc, shift, qmax, rev_n;
    // rows
    int i, sh0, sh1, shm, gmax, te, n_b, last_b_sc, last_b_i;   // sh*: bit offsets of H(i-1,.), H(i,.) and the kept row in the cell word
    int jc, rf, rf2, rimax, rdiag, rcnt, t_now, t_next;
    uint64_t rowpack;
    uint64_t *b; int cap_b;
    Row W; int cap_cells;          // cells of the tile; code word of cell q: W.at(cap_cells + (q >> 3)), four bits each
    SwResult fwd, out;
    int err;

    BSB_HD int qcode(int q) const                     // query of the current pass
    {
        const int p = pass ? fwd.qe - q : q;          // reverse pass: the prefix [0, qe] read backwards
        const int c = is_rev ? ms[l_ms - 1 - p] : ms[p];
        return is_rev ? (c < 4 ? 3 - c : 4) : c;
    }
    BSB_HD int tbase(int k) const                     // target of the current pass
    {
        const int p = pass ? (k < rev_n ? rev_n - 1 - k : k) : k;
        return ref_base(pac, l_pac, rb + p);
    }

    BSB_HD void begin(const Opt &opt, const IndexView &ix, const RescueJob &jb, const uint8_t *mate_seq)
    {
        pac = ix.pac; l_pac = ix.l_pac; rb = jb.rb; ms = mate_seq; l_ms = jb.qlen; is_rev = jb.is_rev; tlen = jb.tlen;
        min_score = opt.min_seed_len * opt.a;
        int sh = 127, md = 0;
        for (int a = 0; a < 25; ++a) { if (opt.mat[a] < (int8_t)sh) sh = opt.mat[a]; if (opt.mat[a] > (int8_t)md) md = opt.mat[a]; }
        qmax = md; shift = (256 - sh) & 0xff;
        out.score = 0; out.te = out.qe = out.score2 = out.te2 = out.tb = out.qb = -1;
        fwd = out;
        err = 0;
        start_pass(0, l_ms, min_score, 0x10000);
    }

    BSB_HD void start_pass(int which, int n_query, int minsc_, int endsc_)
    {
        pass = which; qlen = n_query; slen = (qlen + 15) >> 4; L = slen << 4;
        minsc = minsc_; endsc = endsc_;
        rev_n = fwd.te + 1;
        i = 0; sh0 = 0; sh1 = 8; shm = 24; gmax = 0; te = -1; n_b = 0; last_b_sc = 0; last_b_i = -2;
        jc = 0;
        state = L > cap_cells ? PASS_END : INIT;      // (a query longer than the tile: reported as an error)
        if (L > cap_cells) err = ERR_SCRATCH_OVERFLOW;
    }

    // zeroes up to C cells of the tile and packs their query codes (pad cells: code 5, which scores 0 against everything)
    BSB_HD void init_step(int C)
    {
        int q = jc;
        const int stop_at = q + C < L ? q + C : L;
        for (; q < stop_at; q += 8) {                 // C and L are multiples of 8
            uint32_t codes = 0;
            for (int k = 0; k < 8; ++k) {
                *W.at(q + k) = 0;
                codes |= (uint32_t)(q + k < qlen ? qcode(q + k) : 5) << (4 * k);
            }
            *W.at(cap_cells + (q >> 3)) = codes;
        }
        jc = q;
        if (q >= L) {
            jc = -1;
            if (tlen <= 0) { state = PASS_END; return; }
            t_next = tbase(0);
            state = ROWS;
        }
    }

    // up to C cells of row i; closes the row when its last cell is done
    BSB_HD void step(const Opt &opt, int C)
    {
        const int e_del = opt.e_del, e_ins = opt.e_ins, oe_del = opt.o_del + opt.e_del, oe_ins = opt.o_ins + opt.e_ins;
        if (jc < 0) {                                  // open row i
            t_now = t_next;
            if (i + 1 < tlen) t_next = tbase(i + 1);   // the next row's base is on its way while this row is filled
            const int8_t *row = opt.mat + t_now * 5;
            rowpack = (uint64_t)(uint8_t)row[0] | (uint64_t)(uint8_t)row[1] << 8 | (uint64_t)(uint8_t)row[2] << 16 |
                      (uint64_t)(uint8_t)row[3] << 24 | (uint64_t)(uint8_t)row[4] << 32;   // byte 5 = 0: the pad cells
            rf = rf2 = rimax = rdiag = 0; rcnt = 0;
            jc = 0;
        }
        {
            int q = jc, f = rf, f2 = rf2, imax = rimax, diag = rdiag, cnt = rcnt;
            const int stop_at = q + C < L ? q + C : L;
            const int s0 = sh0, s1 = sh1;
            const uint32_t keep = ~(0xffu << s1) & ~0x00ff0000u;   // everything but H(i, .) and E, which this row writes
            for (; q < stop_at; q += 8) {
                const uint32_t codes = *W.at(cap_cells + (q >> 3));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t *wp = W.at(q + k);
                    const uint32_t w = *wp;
                    if (cnt == 0) { f = 0; cnt = slen; }        // head of a stripe: the first sweep's F restarts
                    --cnt;
                    int hh = diag;
                    diag = (int)(w >> s0 & 0xffu);              // H(i-1, q): the diagonal source of the next cell
                    const int sc = (int)(int8_t)(rowpack >> (8 * (codes >> (4 * k) & 15u)));
                    hh = hh + sc + shift; hh = hh < 255 ? hh : 255; hh -= shift; hh = hh > 0 ? hh : 0;
                    const int e = (int)(w >> 16 & 0xffu);
                    const int h = bsb_max3(hh, e, f);           // first sweep
                    imax = imax > h ? imax : h;
                    const int e2 = bsb_addmax_relu(e, -e_del, h - oe_del);
                    f = bsb_addmax_relu(f, -e_ins, h - oe_ins);
                    const int hf = h > f2 ? h : f2;             // lazy-F repair: what the row finally holds
                    f2 = bsb_addmax_relu(f2, -e_ins, hf - oe_ins);
                    *wp = (w & keep) | (uint32_t)hf << s1 | (uint32_t)e2 << 16;
                }
            }
            jc = q; rf = f; rf2 = f2; rimax = imax; rdiag = diag; rcnt = cnt;
            if (q < L) return;
        }
        // close row i
        jc = -1;
        const int imax = rimax;
        if (imax >= minsc) {
            if (n_b == 0 || last_b_i + 1 != i) {
                if (n_b >= cap_b) { err = ERR_SCRATCH_OVERFLOW; state = PASS_END; return; }
                b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
                last_b_sc = imax; last_b_i = i;
            } else if (last_b_sc < imax) {
                b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
                last_b_sc = imax; last_b_i = i;
            }                      // (an entry that is not raised keeps its row: the next row then starts a new entry, like the reference)
        }
        bool stop = false;
        if (imax > gmax) {
            gmax = imax; te = i;
            shm = sh1;                                 // this row is the one to keep (Hmax of the reference: a copy there)
            if (gmax + shift >= 255 || gmax >= endsc) stop = true;
        }
        if (!stop) {
            sh0 = sh1;                                 // the row just filled is the previous row of the next one,
            sh1 = (sh0 != 0 && shm != 0) ? 0 : (sh0 != 8 && shm != 8) ? 8 : 24;   // which goes into a byte that is neither it nor the kept row
            if (++i >= tlen) stop = true;
        }
        if (stop) state = PASS_END;
    }

    // the end of a pass: the forward result, the decision about the reverse pass, the final result. true: job finished
    BSB_HD bool end_pass()
    {
        SwResult r;
        r.score = gmax + shift < 255 ? gmax : 255;
        r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
        if (err) { out = r; state = IDLE; return true; }
        if (r.score != 255) {
            int mx = -1;
            for (int q = 0; q < L; ++q) { const int t = (int)(*W.at(q) >> shm & 0xffu); if (t > mx) { mx = t; r.qe = q; } }   // smallest position holding the maximum
            if (n_b) {
                int k = (r.score + qmax - 1) / qmax;
                const int low = te - k, high = te + k;
                for (k = 0; k < n_b; ++k) {
                    const int e = (int32_t)b[k];
                    if ((e < low || e > high) && (int)(b[k] >> 32) > r.score2) { r.score2 = (int)(b[k] >> 32); r.te2 = e; }
                }
            }
        }
        if (pass == 0) {
            fwd = r;
            if (r.score < min_score) { out = r; state = IDLE; return true; }
            if (r.score == 255 || r.qe < 0) { err = ERR_SCRATCH_OVERFLOW; out = r; state = IDLE; return true; }   // saturated: not this kernel's case
            start_pass(1, r.qe + 1, 0x10000, r.score);
            return false;
        }
        out = fwd;
        if (fwd.score == r.score) { out.tb = fwd.te - r.te; out.qb = fwd.qe - r.qe; }
        state = IDLE;
        return true;
    }
};

} // namespace bsb
