/* bsb_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the primitives of the `bsbolt Align` hot path, written from
 * the reference's algorithm (NuttyLogic/BSBolt v1.6.0, bsbolt/External/BWA/). It exists so that the
 * CUDA kernels can be checked stage by stage on seeded inputs; it is itself pinned against the real
 * reference functions (oracle/_ref/libbwa_ref.so: bwt_occ4, bwt_extend, bwt_smem1, bwt_sa, ksw_extend2,
 * ksw_global2) by tests/test_oracle.py and, end to end, by the golden SAM files under tests/golden/.
 * Parity status: PINNED (reference outputs generated in the build container; see tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include "bsb_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- read conversion: bsConversion (bs_helpers.cpp:18-31) + nst_nt4_table (bntseq.c:48-65) ---- */
static uint8_t nt4(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case '-': return 5;
    default: return 4;
    }
}

void bso_convert(const char *read, int len, int pattern, uint8_t *seq, uint8_t *oseq)
{
    /* kseq2bseq1 (bwa.c:44-71): oseq keeps the original bases as codes, seq is converted first (upper
     * case only) and coded later by mem_align1_core (bwamem.c:1165-1166) */
    int i;
    for (i = 0; i < len; ++i) {
        char c = read[i];
        oseq[i] = nt4((unsigned char)c);
        if (pattern ? c == 'G' : c == 'C') c = pattern ? 'A' : 'T';
        seq[i] = nt4((unsigned char)c);
    }
}

/* ---- occurrence counts: bwt_occ4 (bwt.c:169-186) with the byte table of bwt_gen_cnt_table (bwt.c:42-51) ---- */
static uint32_t cnt_table[256];
static int cnt_ready = 0;
static void cnt_init(void)
{
    int i, j;
    for (i = 0; i != 256; ++i) {
        uint32_t x = 0;
        for (j = 0; j != 4; ++j)
            x |= (uint32_t)(((i & 3) == j) + ((i >> 2 & 3) == j) + ((i >> 4 & 3) == j) + (i >> 6 == j)) << (j << 3);
        cnt_table[i] = x;
    }
    cnt_ready = 1;
}
static uint32_t word_counts(uint32_t b)
{
    return cnt_table[b & 0xff] + cnt_table[b >> 8 & 0xff] + cnt_table[b >> 16 & 0xff] + cnt_table[b >> 24];
}

void bso_occ4(const bso_index_t *ix, uint64_t k, uint64_t cnt[4])
{
    const uint32_t *p, *end;
    uint64_t x = 0;
    uint32_t tmp;
    if (!cnt_ready) cnt_init();
    if (k == (uint64_t)-1) { memset(cnt, 0, 4 * sizeof(uint64_t)); return; }
    k -= (k >= ix->primary);                      /* '$' is not stored */
    p = ix->bwt + ((k >> 7) << 4);                /* 64-byte block of 128 symbols */
    memcpy(cnt, p, 4 * sizeof(uint64_t));
    p += 8;
    end = p + ((k >> 4) - ((k & ~(uint64_t)0x7f) >> 4));
    for (; p < end; ++p) x += word_counts(*p);
    tmp = *p & ~((1U << ((~k & 15) << 1)) - 1);
    x += word_counts(tmp) - (~k & 15);            /* the masked-out symbols were counted as 'A' */
    cnt[0] += x & 0xff; cnt[1] += x >> 8 & 0xff; cnt[2] += x >> 16 & 0xff; cnt[3] += x >> 24;
}

/* ---- bi-directional extension: bwt_extend (bwt.c:262-275) ---- */
void bso_extend(const bso_index_t *ix, const bso_intv_t *ik, bso_intv_t ok[4], int is_back)
{
    uint64_t tk[4], tl[4];
    int i, nb = !is_back;
    bso_occ4(ix, ik->x[nb] - 1, tk);
    bso_occ4(ix, ik->x[nb] - 1 + ik->x[2], tl);
    for (i = 0; i != 4; ++i) {
        ok[i].x[nb] = ix->L2[i] + 1 + tk[i];
        ok[i].x[2] = tl[i] - tk[i];
    }
    ok[3].x[is_back] = ik->x[is_back] + (ik->x[nb] <= ix->primary && ik->x[nb] + ik->x[2] - 1 >= ix->primary);
    ok[2].x[is_back] = ok[3].x[is_back] + ok[3].x[2];
    ok[1].x[is_back] = ok[2].x[is_back] + ok[2].x[2];
    ok[0].x[is_back] = ok[1].x[is_back] + ok[1].x[2];
}

static void set_intv(const bso_index_t *ix, int c, bso_intv_t *ik)
{   /* bwt_set_intv (bwt.h:80) */
    ik->x[0] = ix->L2[c] + 1; ik->x[2] = ix->L2[c + 1] - ix->L2[c]; ik->x[1] = ix->L2[3 - c] + 1; ik->info = 0;
}

static void reverse_list(bso_intv_t *a, int n)
{
    int j;
    for (j = 0; j < n >> 1; ++j) { bso_intv_t t = a[n - 1 - j]; a[n - 1 - j] = a[j]; a[j] = t; }
}

/* ---- SMEMs through position x: bwt_smem1a with max_intv = 0 (bwt.c:289-351) ----
 * returns the next x, or -1 when `cap` is too small */
int bso_smem(const bso_index_t *ix, int len, const uint8_t *q, int x, int min_intv, bso_intv_t *mem, int *n_mem, int cap)
{
    int i, j, c, ret, np = 0, nc = 0, nm = 0;
    bso_intv_t ik, ok[4], *prev, *curr, *swap;
    *n_mem = 0;
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    prev = (bso_intv_t *)malloc(sizeof(bso_intv_t) * (size_t)(len + 1));
    curr = (bso_intv_t *)malloc(sizeof(bso_intv_t) * (size_t)(len + 1));
    set_intv(ix, q[x], &ik);
    ik.info = (uint64_t)(x + 1);
    for (i = x + 1; i < len; ++i) {                 /* forward: remember every interval-size change */
        if (q[i] < 4) {
            c = 3 - q[i];
            bso_extend(ix, &ik, ok, 0);
            if (ok[c].x[2] != ik.x[2]) {
                curr[nc++] = ik;
                if (ok[c].x[2] < (uint64_t)min_intv) break;
            }
            ik = ok[c]; ik.info = (uint64_t)(i + 1);
        } else { curr[nc++] = ik; break; }
    }
    if (i == len) curr[nc++] = ik;
    reverse_list(curr, nc);
    ret = (int)curr[0].info;
    swap = curr; curr = prev; prev = swap; np = nc;
    for (i = x - 1; i >= -1; --i) {                 /* backward: a match is maximal when it cannot grow */
        c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
        for (j = 0, nc = 0; j < np; ++j) {
            bso_intv_t *p = &prev[j];
            if (c >= 0) bso_extend(ix, p, ok, 1);
            if (c < 0 || ok[c].x[2] < (uint64_t)min_intv) {
                if (nc == 0) {
                    if (nm == 0 || (uint64_t)(i + 1) < mem[nm - 1].info >> 32) {
                        if (nm >= cap) { free(prev); free(curr); return -1; }
                        ik = *p; ik.info |= (uint64_t)(i + 1) << 32;
                        mem[nm++] = ik;
                    }
                }
            } else if (nc == 0 || ok[c].x[2] != curr[nc - 1].x[2]) {
                ok[c].info = p->info;
                curr[nc++] = ok[c];
            }
        }
        if (nc == 0) break;
        swap = curr; curr = prev; prev = swap; np = nc;
    }
    reverse_list(mem, nm);
    *n_mem = nm;
    free(prev); free(curr);
    return ret;
}

/* ---- forward-only seed: bwt_seed_strategy1 (bwt.c:358-379) ---- */
int bso_seed_forward(const bso_index_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, bso_intv_t *mem)
{
    int i, c;
    bso_intv_t ik, ok[4];
    memset(mem, 0, sizeof(bso_intv_t));
    if (q[x] > 3) return x + 1;
    set_intv(ix, q[x], &ik);
    for (i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            c = 3 - q[i];
            bso_extend(ix, &ik, ok, 0);
            if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
                *mem = ok[c];
                mem->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
                return i + 1;
            }
            ik = ok[c];
        } else return i + 1;
    }
    return len;
}

/* ---- suffix array value: bwt_sa / bwt_invPsi / bwt_occ (bwt.c:53-59, 86-129) ---- */
static uint64_t occ1(const bso_index_t *ix, uint64_t k, int c)
{
    uint64_t cnt[4];
    if (k == ix->seq_len) return ix->L2[c + 1] - ix->L2[c];
    if (k == (uint64_t)-1) return 0;
    bso_occ4(ix, k, cnt);   /* the 1-symbol rank (bwt_occ) is one component of the 4-symbol rank */
    return cnt[c];
}

uint64_t bso_sa(const bso_index_t *ix, uint64_t k)
{
    uint64_t sa = 0, mask = (uint64_t)ix->sa_intv - 1;
    while (k & mask) {
        uint64_t x = k - (k > ix->primary);
        int c = (int)(ix->bwt[((x >> 7) << 4) + 8 + ((x & 0x7f) >> 4)] >> ((~x & 0xf) << 1) & 3); /* bwt_B0 (bwt.h:78) */
        ++sa;
        k = k == ix->primary ? 0 : ix->L2[c] + occ1(ix, k, c);
    }
    return sa + ix->sa[k / (uint64_t)ix->sa_intv];
}

/* ---- banded extension: ksw_extend2 (ksw.c:380-479) ---- */
int bso_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore_, int *max_off_)
{
    typedef struct { int32_t h, e; } cell_t;
    cell_t *eh = (cell_t *)calloc((size_t)qlen + 1, sizeof(cell_t));
    int i, j, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
    eh[0].h = h0; eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;
    for (i = 0, max = 0; i < 25; ++i) max = max > mat[i] ? max : mat[i];
    max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
    beg = 0; end = qlen;
    for (i = 0; i < tlen; ++i) {
        int t, f = 0, h1, m = 0, mj = -1;
        const int8_t *sc = mat + target[i] * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        for (j = beg; j < end; ++j) {
            cell_t *p = &eh[j];
            int h, M = p->h, e = p->e;
            p->h = h1;
            M = M ? M + sc[query[j]] : 0;      /* a zero cell cannot restart a match (ksw.c:433) */
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            mj = m > h ? mj : j;               /* last column holding the row maximum */
            m = m > h ? m : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            p->e = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
        }
        eh[end].h = h1; eh[end].e = 0;
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            max_off = max_off > abs(mj - i) ? max_off : abs(mj - i);
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
            else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
        }
        for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    free(eh);
    if (qle) *qle = max_j + 1;
    if (tle) *tle = max_i + 1;
    if (gtle) *gtle = max_ie + 1;
    if (gscore_) *gscore_ = gscore;
    if (max_off_) *max_off_ = max_off;
    return max;
}

/* ---- banded global alignment with traceback: ksw_global2 (ksw.c:504-606) ---- */
#define NEG_INF (-0x40000000)
static int push_op(uint32_t *cigar, int *n, int cap, int op, int len)
{
    if (*n == 0 || (uint32_t)op != (cigar[*n - 1] & 0xf)) {
        if (*n >= cap) return -1;
        cigar[(*n)++] = (uint32_t)len << 4 | (uint32_t)op;
    } else cigar[*n - 1] += (uint32_t)len << 4;
    return 0;
}

int bso_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar, int cigar_cap)
{
    typedef struct { int32_t h, e; } cell_t;
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score;
    int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    uint8_t *z = (uint8_t *)malloc((size_t)n_col * (size_t)tlen + 1);
    cell_t *eh = (cell_t *)calloc((size_t)qlen + 1, sizeof(cell_t));
    eh[0].h = 0; eh[0].e = NEG_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = NEG_INF; }
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = NEG_INF;
    for (i = 0; i < tlen; ++i) {
        int32_t f = NEG_INF, h1, beg, end, t;
        const int8_t *sc = mat + target[i] * 5;
        uint8_t *zi = z + (size_t)i * n_col;
        beg = i > w ? i - w : 0;
        end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
        for (j = beg; j < end; ++j) {
            cell_t *p = &eh[j];
            int32_t h, m = p->h, e = p->e;
            uint8_t d;
            p->h = h1;
            m += sc[query[j]];
            d = m >= e ? 0 : 1;                 /* ties prefer the diagonal, then deletion (ksw.c:551-555) */
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del; e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = m - oe_ins; f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = NEG_INF;
    }
    score = eh[qlen].h;
    if (n_cigar && cigar) {
        int n = 0, which = 0, bad = 0;
        i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        while (i >= 0 && k >= 0) {
            which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { bad |= push_op(cigar, &n, cigar_cap, 0, 1); --i; --k; }
            else if (which == 1) { bad |= push_op(cigar, &n, cigar_cap, 2, 1); --i; }
            else { bad |= push_op(cigar, &n, cigar_cap, 1, 1); --k; }
        }
        if (i >= 0) bad |= push_op(cigar, &n, cigar_cap, 2, i + 1);
        if (k >= 0) bad |= push_op(cigar, &n, cigar_cap, 1, k + 1);
        for (i = 0; i < n >> 1; ++i) { uint32_t t = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = t; }
        *n_cigar = bad ? -1 : n;
    }
    free(eh); free(z);
    return score;
}
