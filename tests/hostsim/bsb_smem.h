// bsb_smem.h -- SMEM seeding over the bi-directional FM-index (north_star stage 2).
//
// Restates, for one read, the interval collection of the reference:
//   smem_at()        <- bwt_smem1a          (bwt.c:289-351), max_intv == 0 as used by `mem`
//   seed_forward()   <- bwt_seed_strategy1  (bwt.c:358-379)
//   collect_intv()   <- mem_collect_intv    (bwamem.c:118-166)
// Interval lists live in caller-provided scratch (HBM on the device); `cap` bounds every list and
// an overflow is reported, never truncated silently.
#pragma once
#include "../../bsbolt_b200/csrc/bsb_index.h"

namespace bsb {

struct IntvList {
    Intv *a;
    int n, cap;
    BSB_HD bool push(const Intv &v) { if (n >= cap) return false; a[n++] = v; return true; }
    BSB_HD void reverse() { for (int j = 0; j < n >> 1; ++j) tswap(a[j], a[n - 1 - j]); }
};

// All SMEMs covering query position x. Returns the end of the longest match starting at x.
BSB_HD int smem_at(const IndexView &ix, int len, const uint8_t *q, int x, int min_intv,
                   IntvList &mem, IntvList &t0, IntvList &t1, int *err)
{
    Intv ik, ok[4];
    mem.n = 0;
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    IntvList *prev = &t0, *curr = &t1;
    fm_set_intv(ix, q[x], ik);
    ik.info = x + 1;
    int i;
    curr->n = 0;
    for (i = x + 1; i < len; ++i) { // forward extension, recording every change of interval size
        if (q[i] < 4) {
            int c = 3 - q[i];
            fm_extend(ix, ik, ok, 0);
            if (ok[c].x2 != ik.x2) {
                if (!curr->push(ik)) *err = ERR_INTV_OVERFLOW;
                if (ok[c].x2 < (uint64_t)min_intv) break;
            }
            ik = ok[c]; ik.info = i + 1;
        } else {
            if (!curr->push(ik)) *err = ERR_INTV_OVERFLOW;
            break;
        }
    }
    if (i == len) { if (!curr->push(ik)) *err = ERR_INTV_OVERFLOW; }
    curr->reverse(); // longest match first
    int ret = (int)curr->a[0].info;
    { IntvList *s = curr; curr = prev; prev = s; }

    for (i = x - 1; i >= -1; --i) { // backward extension; keep a match when it cannot grow further
        int c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
        curr->n = 0;
        for (int j = 0; j < prev->n; ++j) {
            const Intv p = prev->a[j];
            if (c >= 0) fm_extend(ix, p, ok, 1);
            if (c < 0 || ok[c].x2 < (uint64_t)min_intv) {
                if (curr->n == 0) {
                    if (mem.n == 0 || (uint64_t)(i + 1) < (mem.a[mem.n - 1].info >> 32)) {
                        ik = p; ik.info |= (uint64_t)(i + 1) << 32;
                        if (!mem.push(ik)) *err = ERR_INTV_OVERFLOW;
                    }
                }
            } else if (curr->n == 0 || ok[c].x2 != curr->a[curr->n - 1].x2) {
                ok[c].info = p.info;
                if (!curr->push(ok[c])) *err = ERR_INTV_OVERFLOW;
            }
        }
        if (curr->n == 0) break;
        { IntvList *s = curr; curr = prev; prev = s; }
    }
    mem.reverse(); // sorted by start coordinate
    return ret;
}

// forward-only seed: stop as soon as the interval is smaller than max_intv and long enough
BSB_HD int seed_forward(const IndexView &ix, int len, const uint8_t *q, int x, int min_len, int max_intv, Intv *mem)
{
    Intv ik, ok[4];
    mem->x0 = mem->x1 = mem->x2 = mem->info = 0;
    if (q[x] > 3) return x + 1;
    fm_set_intv(ix, q[x], ik);
    for (int i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            int c = 3 - q[i];
            fm_extend(ix, ik, ok, 0);
            if (ok[c].x2 < (uint64_t)max_intv && i - x >= min_len) {
                *mem = ok[c];
                mem->info = (uint64_t)x << 32 | (uint32_t)(i + 1);
                return i + 1;
            }
            ik = ok[c];
        } else return i + 1;
    }
    return len;
}

struct LtIntvInfo { BSB_HD bool operator()(const Intv &a, const Intv &b) const { return a.info < b.info; } };

// Three seeding passes + sort. `mem` receives the final list (sorted by info).
BSB_HD void collect_intv(const Opt &opt, const IndexView &ix, int len, const uint8_t *seq,
                         IntvList &mem, IntvList &mem1, IntvList &t0, IntvList &t1, int *err)
{
    int x = 0;
    const int start_width = 1;
    int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
    mem.n = 0;
    while (x < len) { // pass 1: all SMEMs
        if (seq[x] < 4) {
            x = smem_at(ix, len, seq, x, start_width, mem1, t0, t1, err);
            for (int i = 0; i < mem1.n; ++i) {
                const Intv &p = mem1.a[i];
                int slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
                if (slen >= opt.min_seed_len) { if (!mem.push(p)) *err = ERR_INTV_OVERFLOW; }
            }
        } else ++x;
    }
    int old_n = mem.n;
    for (int k = 0; k < old_n; ++k) { // pass 2: re-seed inside long, rare SMEMs
        const Intv p = mem.a[k];
        int start = (int)(p.info >> 32), end = (int32_t)p.info;
        if (end - start < split_len || p.x2 > (uint64_t)opt.split_width) continue;
        smem_at(ix, len, seq, (start + end) >> 1, (int)(p.x2 + 1), mem1, t0, t1, err);
        for (int i = 0; i < mem1.n; ++i) {
            const Intv &m = mem1.a[i];
            if ((int)((uint32_t)m.info - (uint32_t)(m.info >> 32)) >= opt.min_seed_len) { if (!mem.push(m)) *err = ERR_INTV_OVERFLOW; }
        }
    }
    if (opt.max_mem_intv > 0) { // pass 3: LAST-like forward seeds
        x = 0;
        while (x < len) {
            if (seq[x] < 4) {
                Intv m;
                x = seed_forward(ix, len, seq, x, opt.min_seed_len, (int)opt.max_mem_intv, &m);
                if (m.x2 > 0) { if (!mem.push(m)) *err = ERR_INTV_OVERFLOW; }
            } else ++x;
        }
    }
    introsort((long)mem.n, mem.a, LtIntvInfo());
}

} // namespace bsb
