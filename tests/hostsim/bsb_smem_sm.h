// bsb_smem_sm.h -- SMEM seeding as a per-thread state machine with ONE FM-index extension site.
//
// collect_intv() (bsb_smem.h) nests three loops around bwt_extend; with one read per thread the 32 lanes of
// a warp sit in different loops and only ~5 of them execute any given instruction (ncu, profiles/
// r01_ncu_v1_thread_per_read.md). The extension -- two random 64-byte block loads plus the popcounts -- is
// where the time goes, so the same algorithm is restated here as a state machine: each lane advances its own
// control state (cheap, divergent) until it needs an extension, then ALL lanes meet at the single fm_extend()
// call (expensive, converged). Results are identical to collect_intv(): same passes (mem_collect_intv,
// bwamem.c:118-166), same SMEM sweep (bwt_smem1a, bwt.c:289-351), same forward seeds (bwt_seed_strategy1,
// bwt.c:358-379), same final sort.
#pragma once
#include "bsb_smem.h"

#if defined(__CUDA_ARCH__)
#define BSB_SYNCWARP() __syncwarp()
#define BSB_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define BSB_SYNCWARP()
#define BSB_ANY(p) (p)
#endif

namespace bsb {

struct SeedMachine {
    enum Sub { NEXT, FWD_STEP, FWD_GOT, BWD_INIT, BWD_ROW, BWD_CELL, BWD_GOT, SMEM_END, P3_STEP, P3_GOT, FINISHED };
    // inputs
    const Opt *opt; int len; const uint8_t *q;
    Intv *mem_a, *mem1_a, *prev_a, *curr_a;   // raw HBM pointers + counts: nothing here needs an address, so the
    int mem_n, mem1_n, prev_n, curr_n, cap;   // whole machine lives in registers
    // control
    int sub, pass, x, i, j, c, min_intv, ret, k2, old_n, split_len, err;
    Intv ik;
    // request
    Intv req; int req_back, req_c;

    BSB_HD void init(const Opt &o, int len_, const uint8_t *q_, Intv *pmem, Intv *pmem1, Intv *pt0, Intv *pt1, int cap_)
    {
        opt = &o; len = len_; q = q_;
        mem_a = pmem; mem_n = 0; cap = cap_;
        mem1_a = pmem1; mem1_n = 0;
        prev_a = pt0; prev_n = 0; curr_a = pt1; curr_n = 0;
        sub = NEXT; pass = 1; x = 0; err = 0; k2 = 0; old_n = 0;
        split_len = (int)(o.min_seed_len * o.split_factor + .499);
    }
    BSB_HD void push(Intv *a, int &n, const Intv &v) { if (n < cap) a[n++] = v; else err = ERR_INTV_OVERFLOW; }
    BSB_HD void swap_lists() { Intv *t = curr_a; curr_a = prev_a; prev_a = t; int k = curr_n; curr_n = prev_n; prev_n = k; }
    static BSB_HD void reverse(Intv *a, int n) { for (int j = 0; j < n >> 1; ++j) tswap(a[j], a[n - 1 - j]); }

    BSB_HD void start_smem(const IndexView &ix, int x_, int min_intv_)
    {   // head of smem_at(): caller guarantees q[x_] < 4
        x = x_; min_intv = min_intv_ < 1 ? 1 : min_intv_;
        mem1_n = 0;
        fm_set_intv(ix, q[x], ik);
        ik.info = (uint64_t)(x + 1);
        curr_n = 0;
        i = x + 1;
        sub = FWD_STEP;
    }

    // Advances until an extension is required (returns true, request in req/req_back) or the read is done.
    BSB_HD bool advance(const IndexView &ix)
    {
        for (;;) {
            switch (sub) {
            case NEXT:
                if (pass == 1) {
                    while (x < len && q[x] >= 4) ++x;
                    if (x >= len) { pass = 2; old_n = mem_n; k2 = 0; break; }
                    start_smem(ix, x, 1);
                } else if (pass == 2) {
                    bool started = false;
                    while (k2 < old_n) {
                        const Intv p = mem_a[k2];
                        int start = (int)(p.info >> 32), end = (int32_t)p.info;
                        if (end - start < split_len || p.x2 > (uint64_t)opt->split_width) { ++k2; continue; }
                        int xm = (start + end) >> 1;
                        if (q[xm] > 3) { mem1_n = 0; ++k2; continue; } // smem_at() returns at once on an ambiguous base
                        start_smem(ix, xm, (int)(p.x2 + 1));
                        started = true;
                        break;
                    }
                    if (!started) { pass = 3; x = 0; sub = opt->max_mem_intv > 0 ? P3_STEP : FINISHED; i = -1; }
                } else sub = FINISHED;
                break;
            case FWD_STEP:
                if (i >= len) { push(curr_a, curr_n, ik); sub = BWD_INIT; }
                else if (q[i] < 4) { req = ik; req_back = 0; req_c = 3 - q[i]; sub = FWD_GOT; return true; }
                else { push(curr_a, curr_n, ik); sub = BWD_INIT; }
                break;
            case BWD_INIT:
                reverse(curr_a, curr_n);
                ret = (int)curr_a[0].info;
                swap_lists();
                i = x - 1;
                sub = BWD_ROW;
                break;
            case BWD_ROW:
                if (i < -1) { sub = SMEM_END; break; }
                c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
                curr_n = 0; j = 0;
                sub = BWD_CELL;
                break;
            case BWD_CELL:
                if (j >= prev_n) {
                    if (curr_n == 0) { sub = SMEM_END; break; }
                    swap_lists();
                    --i;
                    sub = BWD_ROW;
                } else if (c >= 0) { req = prev_a[j]; req_back = 1; req_c = c; sub = BWD_GOT; return true; }
                else {
                    if (curr_n == 0) {
                        if (mem1_n == 0 || (uint64_t)(i + 1) < (mem1_a[mem1_n - 1].info >> 32)) {
                            Intv t = prev_a[j]; t.info |= (uint64_t)(i + 1) << 32;
                            push(mem1_a, mem1_n, t);
                        }
                    }
                    ++j;
                }
                break;
            case SMEM_END:
                reverse(mem1_a, mem1_n);
                for (int t = 0; t < mem1_n; ++t) {
                    const Intv p = mem1_a[t];
                    if ((int)((uint32_t)p.info - (uint32_t)(p.info >> 32)) >= opt->min_seed_len) push(mem_a, mem_n, p);
                }
                if (pass == 1) x = ret; else ++k2;
                sub = NEXT;
                break;
            case P3_STEP:
                if (i < 0) { // pick the next start
                    while (x < len && q[x] >= 4) ++x;
                    if (x >= len) { sub = FINISHED; break; }
                    fm_set_intv(ix, q[x], ik);
                    i = x + 1;
                }
                if (i >= len) { x = len; i = -1; break; }           // seed_forward() returns len
                if (q[i] < 4) { req = ik; req_back = 0; req_c = 3 - q[i]; sub = P3_GOT; return true; }
                x = i + 1; i = -1;                                   // ambiguous base: return i + 1
                break;
            case FINISHED:
                return false;
            default:
                return false;
            }
        }
    }

    BSB_HD void consume(const Intv &o)   // o = the extension of req by symbol req_c
    {
        if (sub == FWD_GOT) {
            if (o.x2 != ik.x2) {
                push(curr_a, curr_n, ik);
                if (o.x2 < (uint64_t)min_intv) { sub = BWD_INIT; return; }
            }
            ik = o; ik.info = (uint64_t)(i + 1);
            ++i;
            sub = FWD_STEP;
        } else if (sub == BWD_GOT) {
            const Intv p = req; // == prev_a[j]
            if (o.x2 < (uint64_t)min_intv) {
                if (curr_n == 0) {
                    if (mem1_n == 0 || (uint64_t)(i + 1) < (mem1_a[mem1_n - 1].info >> 32)) {
                        Intv t = p; t.info |= (uint64_t)(i + 1) << 32;
                        push(mem1_a, mem1_n, t);
                    }
                }
            } else if (curr_n == 0 || o.x2 != curr_a[curr_n - 1].x2) {
                Intv t = o; t.info = p.info;
                push(curr_a, curr_n, t);
            }
            ++j;
            sub = BWD_CELL;
        } else if (sub == P3_GOT) {
            if (o.x2 < opt->max_mem_intv && i - x >= opt->min_seed_len) {
                Intv m = o;
                m.info = (uint64_t)x << 32 | (uint32_t)(i + 1);
                if (m.x2 > 0) push(mem_a, mem_n, m);
                x = i + 1; i = -1;
            } else { ik = o; ++i; }
            sub = P3_STEP;
        }
    }
};

// Drop-in for collect_intv(); on the device all lanes of the warp must call it together.
BSB_HD void collect_intv_sm(const Opt &opt, const IndexView &ix, int len, const uint8_t *seq,
                            IntvList &mem, IntvList &mem1, IntvList &t0, IntvList &t1, int *err, bool active)
{
    SeedMachine sm;
    sm.init(opt, len, seq, mem.a, mem1.a, t0.a, t1.a, mem.cap);
    if (!active) sm.sub = SeedMachine::FINISHED;
    for (;;) {
        bool need = sm.advance(ix);
        BSB_SYNCWARP();
        if (!BSB_ANY(need)) break;
        if (need) {
            const Intv o = fm_extend_sel(ix, sm.req, sm.req_c, sm.req_back);
            sm.consume(o);
        }
    }
    mem.n = sm.mem_n;
    if (sm.err) *err = sm.err;
    if (active) introsort((long)mem.n, mem.a, LtIntvInfo());
}

} // namespace bsb
