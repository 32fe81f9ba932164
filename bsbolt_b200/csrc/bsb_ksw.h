// bsb_ksw.h -- integer dynamic-programming primitives, one sequence pair per call (scalar form).
//
//   (ksw_extend2, ksw.c:380-479: ExtLane::step of bsb_extlane.h and sw_extend_warp of bsb_warp.cuh; the plain scalar form of
//    the CPU harness is tests/hostsim/scalar_stages.h)
//   nw_global()  <- ksw_global2  (ksw.c:504-606)  banded global alignment + traceback -> CIGAR
//   sw_local()   <- ksw_u8 / ksw_i16 / ksw_align2 (ksw.c:111-365) local alignment used by mate rescue
//
// These are the exact-semantics building blocks: every tie rule, band update and the persistent
// (h,e) row state of the reference is kept, because score, end points and max_off feed decisions
// downstream. Sequences are read through small accessor functors so that reference bases come
// straight from the 2-bit pac in HBM (no per-call materialisation).
// Attribution: restates klib's ksw.c (ksw_global2, ksw_u8, ksw_i16, ksw_align2; MIT, Attractive Chaos) cell for cell. See NOTICE.md.
#pragma once
#include "bsb_index.h"

namespace bsb {

struct QrySeq {           // query bases: base[i*dir]
    const uint8_t *base; int dir;
    BSB_HD int operator()(int i) const { return base[(long)i * dir]; }
};
struct RefSeq {           // reference bases in the doubled coordinate space: start + i*dir
    const uint8_t *pac; int64_t l_pac, start; int dir;
    BSB_HD int operator()(int i) const { return ref_base(pac, l_pac, start + (int64_t)i * dir); }
};

struct ExtResult { int score, qle, tle, gtle, gscore, max_off; };

// Early end of the extension loop (both forms of sw_extend). ksw_extend2 keeps filling rows until the row maximum is 0,
// z-drop fires or the target window ends -- for a query that has been consumed that is up to max_gap more full-width
// rows whose only possible effects are a larger `max`, or a to-end score >= gscore. Once the band has reached the end of
// the query (end == qlen: every cell a later row can read was written by this row, nothing stale) define
//     phi = max_j ( H[j] > 0 ? H[j] + amax*(qlen-j) : 0 ,  E[j] > 0 ? E[j] + amax*(qlen-1-j) : 0 )
// over the stored row (H[j] = H(i,j-1), the diagonal source of column j; E[j] = E(i+1,j); amax = largest matrix entry).
// Every cell of every later row satisfies h(i',j') + amax*(qlen-1-j') <= phi: M gains at most amax per column, E and F
// only lose, zero cells stay zero (M = H ? H + s : 0), and the first-column source h0 - o_del - e_del*(i'+1) only shrinks.
// So when phi <= max and phi < gscore no later row can change max/max_i/max_j/max_off (they need m > max) or
// gscore/max_ie (they need h1 >= gscore): the loop may stop with identical results.
BSB_HD bool ext_rows_exhausted(int phi, int max, int gscore) { return phi <= max && phi < gscore; }

#define BSB_MINUS_INF (-0x40000000)

struct CigarBuf {
    uint32_t *a; int n, cap;
    BSB_HD bool push(int op, int len)
    {
        if (n == 0 || (uint32_t)op != (a[n - 1] & 0xf)) {
            if (n >= cap) return false;
            a[n++] = (uint32_t)len << 4 | (uint32_t)op;
        } else a[n - 1] += (uint32_t)len << 4;
        return true;
    }
};

// eh: 2*(qlen+1) ints; z: n_col*tlen bytes (only when cig != nullptr). Returns the global score.
template <class Q, class T>
BSB_HD int nw_global(int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                     int o_del, int e_del, int o_ins, int e_ins, int w, int32_t *eh, uint8_t *z, CigarBuf *cig, int *err)
{
    int i, j, oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int32_t *H = eh, *E = eh + (qlen + 1);
    if (cig) cig->n = 0;
    H[0] = 0; E[0] = BSB_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = BSB_MINUS_INF; }
    for (; j <= qlen; ++j) H[j] = E[j] = BSB_MINUS_INF;
    for (i = 0; i < tlen; ++i) {
        int32_t f = BSB_MINUS_INF, h1, beg, end, t;
        const int8_t *row = mat + target(i) * 5;
        beg = i > w ? i - w : 0;
        end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : BSB_MINUS_INF;
        uint8_t *zi = cig ? z + (long)i * n_col : nullptr;
        for (j = beg; j < end; ++j) {
            int32_t h, m = H[j], e = E[j];
            uint8_t d;
            H[j] = h1;
            m += row[query(j)];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            E[j] = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            if (zi) zi[j - beg] = d;
        }
        H[end] = h1; E[end] = BSB_MINUS_INF;
    }
    int score = H[qlen];
    if (cig) {
        int which = 0, k;
        i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        bool ok = true;
        while (i >= 0 && k >= 0) {
            which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { ok &= cig->push(0, 1); --i; --k; }
            else if (which == 1) { ok &= cig->push(2, 1); --i; }
            else { ok &= cig->push(1, 1); --k; }
        }
        if (i >= 0) ok &= cig->push(2, i + 1);
        if (k >= 0) ok &= cig->push(1, k + 1);
        if (!ok) *err = ERR_CIGAR_OVERFLOW;
        for (i = 0; i < cig->n >> 1; ++i) tswap(cig->a[i], cig->a[cig->n - 1 - i]);
    }
    return score;
}

// ---------------------------------------------------------------------------------------------
// Local alignment with the exact semantics of the reference's striped SSE2 kernels.
//
// ksw_u8/ksw_i16 lay the query out in p = 16 (bytes) or 8 (words) stripes of slen cells. Two
// properties of that layout are result-visible and are reproduced cell by cell:
//   * E(i+1,j) is derived from the H(i,j) of the first sweep, in which F restarts from zero at
//     the head of every stripe (ksw.c:172-194); the lazy-F loop afterwards only repairs H.
//   * the padded cells q >= qlen score 0 against everything and take part in the row maximum, so
//     a high score at the end of the query is carried into later rows of the sub-optimal list b[].
// Arithmetic is saturating at 0 (and at 255 for bytes) like the intrinsics.
// ---------------------------------------------------------------------------------------------
struct SwResult { int score, te, qe, score2, te2, tb, qb; };

enum : int { SW_XBYTE = 0x10000, SW_XSTOP = 0x20000, SW_XSUBO = 0x40000, SW_XSTART = 0x80000 };

struct SwScratch {
    int32_t *H0, *H1, *E, *Hmax; // each cap cells
    uint64_t *b;                 // sub-optimal list, cap_b entries
    int cap, cap_b;
};

BSB_HD int sat0(int x) { return x > 0 ? x : 0; }

template <class Q, class T>
BSB_HD SwResult sw_striped(int size, int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                           int o_del, int e_del, int o_ins, int e_ins, int xtra, SwScratch &ws, int *err)
{
    SwResult r = {0, -1, -1, -1, -1, -1, -1};
    const int p = size == 1 ? 16 : 8;
    const int slen = (qlen + p - 1) / p, L = slen * p;
    int shift = 127, mdiff = 0, qmax;
    for (int a = 0; a < 25; ++a) {
        if (mat[a] < (int8_t)shift) shift = mat[a];
        if (mat[a] > (int8_t)mdiff) mdiff = mat[a];
    }
    qmax = mdiff;
    shift = (256 - shift) & 0xff;
    if (L > ws.cap) { *err = ERR_SCRATCH_OVERFLOW; return r; }
    int minsc = (xtra & SW_XSUBO) ? xtra & 0xffff : 0x10000;
    int endsc = (xtra & SW_XSTOP) ? xtra & 0xffff : 0x10000;
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int32_t *H0 = ws.H0, *H1 = ws.H1, *E = ws.E, *Hmax = ws.Hmax;
    int n_b = 0, te = -1, gmax = 0;
    for (int j = 0; j < L; ++j) E[j] = H0[j] = Hmax[j] = 0;
    for (int i = 0; i < tlen; ++i) {
        const int8_t *row = mat + target(i) * 5;
        int f = 0, imax = 0;
        for (int q = 0; q < L; ++q) { // first sweep: F restarts at every stripe head
            if (q % slen == 0) f = 0;
            int h = q ? H0[q - 1] : 0;
            int s = q < qlen ? row[query(q)] : 0;
            if (size == 1) { h = h + s + shift; if (h > 255) h = 255; h = sat0(h - shift); }
            else { h = h + s; if (h > 32767) h = 32767; }
            int e = E[q];
            h = h > e ? h : e;
            h = h > f ? h : f;
            imax = imax > h ? imax : h;
            H1[q] = h;
            e = sat0(e - e_del);
            int t = sat0(h - oe_del);
            E[q] = e > t ? e : t;
            f = sat0(f - e_ins);
            t = sat0(h - oe_ins);
            f = f > t ? f : t;
        }
        f = 0; // lazy-F repair: F carried across stripe boundaries, H only
        for (int q = 0; q < L; ++q) {
            int h = H1[q];
            if (f > h) h = H1[q] = f;
            int t = sat0(h - oe_ins);
            f = sat0(f - e_ins);
            f = f > t ? f : t;
        }
        if (imax >= minsc) {
            if (n_b == 0 || (int32_t)ws.b[n_b - 1] + 1 != i) {
                if (n_b >= ws.cap_b) { *err = ERR_SCRATCH_OVERFLOW; return r; }
                ws.b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
            } else if ((int)(ws.b[n_b - 1] >> 32) < imax) ws.b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
        }
        if (imax > gmax) {
            gmax = imax; te = i;
            for (int q = 0; q < L; ++q) Hmax[q] = H1[q];
            if (size == 1) { if (gmax + shift >= 255 || gmax >= endsc) break; }
            else if (gmax >= endsc) break;
        }
        int32_t *S = H1; H1 = H0; H0 = S;
    }
    r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (size != 1 || r.score != 255) {
        int max = -1;
        if (size != 1) r.qe = -1;
        for (int q = 0; q < L; ++q) { // smallest query position holding the maximum
            int t = Hmax[q];
            if (t > max) { max = t; r.qe = q; }
        }
        if (n_b) {
            int i = (r.score + qmax - 1) / qmax;
            int low = te - i, high = te + i;
            for (i = 0; i < n_b; ++i) {
                int e = (int32_t)ws.b[i];
                if ((e < low || e > high) && (int)(ws.b[i] >> 32) > r.score2) { r.score2 = (int)(ws.b[i] >> 32); r.te2 = e; }
            }
        }
    }
    return r;
}

// ksw_align2 (ksw.c:343-365): forward pass, then a reverse pass over the reversed prefixes to
// locate the start. query/target are forward accessors.
template <class Q, class T>
BSB_HD SwResult sw_local(int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                         int o_del, int e_del, int o_ins, int e_ins, int xtra, SwScratch &ws, int *err)
{
    int size = (xtra & SW_XBYTE) ? 1 : 2;
    SwResult r = sw_striped(size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, ws, err);
    if ((xtra & SW_XSTART) == 0 || ((xtra & SW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    struct RevQ { const Q &q; int n; BSB_HD int operator()(int i) const { return q(n - 1 - i); } };
    struct RevT { const T &t; int n; BSB_HD int operator()(int i) const { return i < n ? t(n - 1 - i) : t(i); } };
    RevQ rq = {query, r.qe + 1};
    RevT rt = {target, r.te + 1};
    SwResult rr = sw_striped(size, r.qe + 1, rq, tlen, rt, mat, o_del, e_del, o_ins, e_ins, SW_XSTOP | r.score, ws, err);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}

} // namespace bsb
