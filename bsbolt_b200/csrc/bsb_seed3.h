// bsb_seed3.h -- SMEM seeding, third form: two independent work items per read, one in-place interval list.
//
// Same results as collect_intv() (bsb_smem.h), i.e. mem_collect_intv (bwamem.c:118-166) over bwt_smem1a
// (bwt.c:289-351) and bwt_seed_strategy1 (bwt.c:358-379); what changed is how the work is laid out:
//
//  * item A of a read = pass 1 (all SMEMs) followed by pass 2 (re-seeding inside long, rare SMEMs);
//    item B = pass 3 (forward-only "LAST-like" seeds). Pass 3 reads nothing the other passes write, so the two
//    items run on different lanes and the list is merged and sorted afterwards (seed3_finish). Interval records
//    with equal `info` describe the same substring of the read and are therefore identical, so ANY sort by
//    `info` gives the list the reference's introsort gives.
//  * the reverse-strand coordinate x[1] of an interval is only needed while it is extended FORWARD. Nothing
//    downstream reads it (mem_chain uses x[0], x[2], info), and the backward sweep of bwt_smem1a extends
//    backward only -- so list entries are (x0, size, end) and the backward sweep rewrites ONE list in place:
//    an entry consumed at rank j produces at most one entry at rank <= j. The forward sweep pushes entries in
//    increasing length; the backward sweep walks them from the top down, so nothing is ever reversed.
//  * every FM-index extension of a lane happens at one call site (the caller's), the lanes of a warp meet
//    there; between extensions a lane touches only its list (shared memory on the device), its base window
//    and registers.
#pragma once
#include "bsb_index.h"

namespace bsb {

// Counts, inside one occ block, of the symbols == c (E) and > c (G) among BWT[block start .. block start + r],
// added to the block's cumulative counts. Same numbers bwt_occ4 (bwt.c:169-186) would give, but only the two
// sums an extension by ONE symbol needs, computed branch-free from the two bit planes of the packed words.
BSB_HD void block_occ_eg(const OccBlock &b, int r, int c, uint64_t &E, uint64_t &G)
{
    const int nsym = r + 1;
    const uint64_t MH = (c & 2) ? ~0ull : 0ull, ML = (c & 1) ? ~0ull : 0ull;
    uint32_t e = 0, g = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t w = (uint64_t)b.w[2 * i] << 32 | b.w[2 * i + 1];
        int k = nsym - 32 * i;
        k = k < 0 ? 0 : k > 32 ? 32 : k;
        uint64_t keep = k == 0 ? 0ull : (~0ull << ((32 - k) << 1));
        keep &= 0x5555555555555555ull;
        const uint64_t hi = w >> 1, lo = w;
        const uint64_t same_hi = ~(hi ^ MH);
        e += popc64(same_hi & ~(lo ^ ML) & keep);
        g += popc64(((hi & ~MH) | (same_hi & lo & ~ML)) & keep);
    }
    uint64_t he = b.cnt[3], hg = 0;
    he = c == 2 ? b.cnt[2] : he; he = c == 1 ? b.cnt[1] : he; he = c == 0 ? b.cnt[0] : he;
    hg += c < 3 ? b.cnt[3] : 0; hg += c < 2 ? b.cnt[2] : 0; hg += c < 1 ? b.cnt[1] : 0;
    E = he + e; G = hg + g;
}

// the extension both sweeps use: interval (xa = the coordinate on the strand being extended, xb = the other one,
// s = size) by symbol c. Returns the new xa-side start, the new xb-side start and the new size.
// Restates bwt_extend (bwt.c:262-275) for one output symbol: with E(p) = occ(c, p) and G(p) = sum of occ(j, p), j > c,
//   na = L2[c] + 1 + E(k),  size = E(l) - E(k),  nb = xb + [sentinel inside] + G(l) - G(k)     (k = xa - 1, l = k + s)
BSB_HD void fm_extend_one(const IndexView &ix, uint64_t xa, uint64_t xb, uint64_t s, int c, uint64_t &na, uint64_t &nb, uint64_t &sz)
{
    const uint64_t k = xa - 1, l = xa - 1 + s;
    const bool kz = k == (uint64_t)-1, lz = l == (uint64_t)-1;
    const uint64_t _k = kz ? 0 : k - (k >= ix.primary), _l = lz ? 0 : l - (l >= ix.primary);
    OccBlock bk, bl;
    load_block(ix.bwt, _k >> 7, bk);
    load_block(ix.bwt, _l >> 7, bl);
    uint64_t ek, gk, el, gl;
    block_occ_eg(bk, (int)(_k & 127), c, ek, gk);
    block_occ_eg(bl, (int)(_l & 127), c, el, gl);
    if (kz) ek = gk = 0;
    if (lz) el = gl = 0;
    na = ix.L2[c] + 1 + ek;
    sz = el - ek;
    nb = xb + (xa <= ix.primary && xa + s - 1 >= ix.primary) + (gl - gk);
}

BSB_HD void fm_extend_any(const IndexView &ix, uint64_t xa, uint64_t xb, uint64_t s, int c, uint64_t &na, uint64_t &nb, uint64_t &sz)
{
    fm_extend_one(ix, xa, xb, s, c, na, nb, sz);
}
BSB_HD void fm_extend_any(const IndexView &ix, uint32_t xa, uint32_t xb, uint32_t s, int c, uint32_t &na, uint32_t &nb, uint32_t &sz)
{
    fm_extend_one32(ix, xa, xb, s, c, na, nb, sz);
}

// Bases: int get(int i) -> code 0..3, >3 ambiguous.
// List: push(p, x0, x2, end) appends rank p during the forward sweep; get(p, nf, ...) / set(p, nf, ...) read and rewrite
// rank p once the forward sweep has ended with nf entries (the device list keeps the top ranks in shared memory); cap().
// U: the coordinate type -- uint32_t when the BWT has fewer than 2^32 symbols (half the registers), else uint64_t.
template <class Bases, class List, class U>
struct Seeder3 {
    enum St { NEXT, FWD, BWD_ROW, BWD_CELL, SMEM_END, P3, W_FWD, W_BWD, W_P3, DONE };
    Bases q; List L;
    Intv *out; int out_cap;          // this read's slice of the interval array: item A fills from the front, item B from the back
    int len;
    int st, pass, x, i, min_intv, ret, c, err;
    U k0, k1, ks;                    // the interval being extended forward (x0, x1, size); it ends at query position i
    int nf, pn, j, cn; U last_x2;    // list: nf entries after the forward sweep; current row = ranks nf-1 .. nf-pn
    U e0, e2; int e_end;             // the list entry whose backward extension is in flight
    int m1_start;                    // bwt_smem1a's "mem->n == 0 || i + 1 < last start" test, without the list (INT_MAX: none yet)
    int k2, old_n, n_out;

    BSB_HD void init(const Opt &o, int len_, Intv *out_, int out_cap_, bool item_b)
    {
        len = len_; out = out_; out_cap = out_cap_; n_out = 0; err = 0;
        x = 0; i = -1; k2 = 0; old_n = 0;
        if (item_b) { pass = 3; st = o.max_mem_intv > 0 ? P3 : DONE; }
        else { pass = 1; st = NEXT; }
    }
    BSB_HD bool done() const { return st == DONE; }

    BSB_HD void emit(U x0, U x2, int start, int end)
    {
        if (n_out >= out_cap) { err = ERR_INTV_OVERFLOW; return; }
        Intv v; v.x0 = x0; v.x1 = 0; v.x2 = x2; v.info = (uint64_t)(uint32_t)start << 32 | (uint32_t)end;
        out[pass == 3 ? out_cap - 1 - n_out : n_out] = v;
        ++n_out;
    }
    BSB_HD void emit_if_new(const Opt &o, U x0, U x2, int start, int end)
    {   // bwt.c:334-338 / 343-344 and the length filter of bwamem.c:130-133
        if (start >= m1_start) return;
        m1_start = start;
        if (end - start >= o.min_seed_len) emit(x0, x2, start, end);
    }
    BSB_HD void push_fwd()
    {
        if (nf >= L.cap()) { err = ERR_INTV_OVERFLOW; return; }
        L.push(nf, k0, ks, i); ++nf; ret = i;
    }
    BSB_HD void begin_bwd() { pn = nf; i = x - 1; st = BWD_ROW; }
    BSB_HD void set_intv(const IndexView &ix, int b)
    {
        k0 = (U)(ix.L2[b] + 1); ks = (U)(ix.L2[b + 1] - ix.L2[b]); k1 = (U)(ix.L2[3 - b] + 1);
    }
    BSB_HD void start_smem(const IndexView &ix, int x_, int min_intv_)
    {   // head of bwt_smem1a; caller guarantees q[x_] < 4
        x = x_; min_intv = min_intv_ < 1 ? 1 : min_intv_;
        m1_start = 0x7fffffff; nf = 0; ret = x + 1;
        set_intv(ix, q.get(x));
        i = x + 1;
        st = FWD;
    }

    // Runs until this lane needs an FM-index extension (true; operands via request()) or has finished (false).
    BSB_HD bool advance(const Opt &o, const IndexView &ix)
    {
        for (;;) {
            switch (st) {
            case NEXT:
                if (pass == 1) {
                    while (x < len && q.get(x) > 3) ++x;
                    if (x >= len) { pass = 2; old_n = n_out; k2 = 0; break; }
                    start_smem(ix, x, 1);
                } else {
                    bool started = false;
                    const int split_len = (int)(o.min_seed_len * o.split_factor + .499);
                    while (k2 < old_n) {
                        const uint64_t info = out[k2].info, occ = out[k2].x2;
                        const int start = (int)(info >> 32), end = (int32_t)info;
                        if (end - start < split_len || occ > (uint64_t)o.split_width) { ++k2; continue; }
                        const int xm = (start + end) >> 1;
                        if (q.get(xm) > 3) { ++k2; continue; }      // bwt_smem1a returns at once on an ambiguous base
                        start_smem(ix, xm, (int)(occ + 1));
                        started = true;
                        break;
                    }
                    if (!started) st = DONE;
                }
                break;
            case FWD:
                if (i < len) {
                    const int b = q.get(i);
                    if (b < 4) { c = 3 - b; st = W_FWD; return true; }
                }
                push_fwd(); begin_bwd();       // read end or ambiguous base: record the interval, turn around
                break;
            case BWD_ROW:
                if (i >= 0) c = q.get(i);
                if (i < 0 || c > 3) {          // nothing can be extended: only the longest entry can be a new SMEM
                    L.get(nf - 1, nf, e0, e2, e_end);
                    emit_if_new(o, e0, e2, i + 1, e_end);
                    st = SMEM_END;
                } else { j = 0; cn = 0; st = BWD_CELL; }
                break;
            case BWD_CELL:
                if (j >= pn) {
                    if (cn == 0) st = SMEM_END;
                    else { pn = cn; --i; st = BWD_ROW; }
                } else { L.get(nf - 1 - j, nf, e0, e2, e_end); st = W_BWD; return true; }
                break;
            case SMEM_END:
                if (pass == 1) x = ret; else ++k2;
                st = NEXT;
                break;
            case P3:
                if (i < 0) {                   // next start
                    while (x < len && q.get(x) > 3) ++x;
                    if (x >= len) { st = DONE; break; }
                    set_intv(ix, q.get(x));
                    i = x + 1;
                }
                if (i >= len) { x = len; i = -1; break; }          // bwt_seed_strategy1 returns len
                {
                    const int b = q.get(i);
                    if (b < 4) { c = 3 - b; st = W_P3; return true; }
                }
                x = i + 1; i = -1;                                  // ambiguous base: restart behind it
                break;
            default:
                return false;
            }
        }
    }

    // operands of the pending extension: xa = coordinate on the strand being extended, xb = the other one, s = size
    BSB_HD void request(U &xa, U &xb, U &s) const
    {
        if (st == W_BWD) { xa = e0; xb = 0; s = e2; } else { xa = k1; xb = k0; s = ks; }
    }

    // Takes the result of the pending extension and runs on to the next one (true) or to the end of the item (false).
    // The common continuations -- next base of a forward sweep, next entry or next row of the backward sweep -- are
    // handled here in straight-line code; everything else goes through the state loop of advance().
    BSB_HD bool step(const Opt &o, const IndexView &ix, U na, U nb, U sz)
    {
        if (st == W_BWD) {
            if (sz < (U)min_intv) {
                if (cn == 0) emit_if_new(o, e0, e2, i + 1, e_end);
            } else if (cn == 0 || sz != last_x2) {
                L.set(nf - 1 - cn, nf, na, sz, e_end); ++cn; last_x2 = sz;
            }
            ++j;
            if (j >= pn && cn > 0 && i > 0) {      // row finished with survivors: next row, if its base can be matched
                const int b = q.get(i - 1);
                if (b < 4) { pn = cn; --i; c = b; j = 0; cn = 0; }
            }
            if (j < pn) { L.get(nf - 1 - j, nf, e0, e2, e_end); return true; }
            st = BWD_CELL;
        } else {
            bool on = true;
            if (st == W_FWD) {
                if (sz != ks) {
                    push_fwd();
                    if (sz < (U)min_intv) { begin_bwd(); on = false; }
                }
            } else if (sz < (U)o.max_mem_intv && i - x >= o.min_seed_len) {    // W_P3: seed found, restart behind it
                if (sz > 0) emit(nb, sz, x, i + 1);
                x = i + 1; i = -1; st = P3; on = false;
            }
            if (on) {
                k1 = na; k0 = nb; ks = sz;
                ++i;
                if (i < len) {
                    const int b = q.get(i);
                    if (b < 4) { c = 3 - b; return true; }
                }
                st = st == W_FWD ? FWD : P3;
            }
        }
        return advance(o, ix);
    }
};

// Merges the two items of a read (front part n_a, back part n_b of `mem`), sorts by info, and runs the tail of
// the seeding stage. Returns the list length or -1 when the two parts collided (capacity too small).
BSB_HD int seed3_merge_sort(Intv *mem, int cap, int n_a, int n_b)
{
    if (n_a + 2 * n_b > cap) return -1;                                // also keeps the copy below clear of its own source
    for (int k = 0; k < n_b; ++k) mem[n_a + k] = mem[cap - 1 - k];
    const int n = n_a + n_b;
    for (int a = 1; a < n; ++a) {
        const Intv v = mem[a];
        int b = a;
        while (b > 0 && mem[b - 1].info > v.info) { mem[b] = mem[b - 1]; --b; }
        mem[b] = v;
    }
    return n;
}

// Plain-memory adapters (CPU unit harness; the device kernel has its own shared-memory list, bsb_cuda.cu)
struct BasesBytes { const uint8_t *p; BSB_HD int get(int i) const { return p[i]; } };
struct ListPlain {
    uint64_t *a0, *a2; int *ae; int n;
    BSB_HD int cap() const { return n; }
    template <class U> BSB_HD void get(int p, int, U &x0, U &x2, int &end) const { x0 = (U)a0[p]; x2 = (U)a2[p]; end = ae[p]; }
    template <class U> BSB_HD void set(int p, int, U x0, U x2, int end) { a0[p] = x0; a2[p] = x2; ae[p] = end; }
    template <class U> BSB_HD void push(int p, U x0, U x2, int end) { a0[p] = x0; a2[p] = x2; ae[p] = end; }
};

// collect_intv() through the two work items, one after the other (what the kernel does on two lanes)
template <class U, class Bases, class List>
BSB_HD int collect_intv_v3(const Opt &opt, const IndexView &ix, int len, Bases q, List L, Intv *mem, int cap, int *err)
{
    int n_part[2] = {0, 0};
    for (int item = 0; item < 2; ++item) {
        Seeder3<Bases, List, U> sm;
        sm.q = q; sm.L = L;
        sm.init(opt, len, mem, cap, item == 1);
        bool need = sm.advance(opt, ix);
        while (need) {
            U xa, xb, s, na, nb, sz;
            sm.request(xa, xb, s);
            fm_extend_any(ix, xa, xb, s, sm.c, na, nb, sz);
            need = sm.step(opt, ix, na, nb, sz);
        }
        if (sm.err) *err = sm.err;
        n_part[item] = sm.n_out;
    }
    if (*err) return 0;
    const int n = seed3_merge_sort(mem, cap, n_part[0], n_part[1]);
    if (n < 0) { *err = ERR_INTV_OVERFLOW; return 0; }
    return n;
}

} // namespace bsb
