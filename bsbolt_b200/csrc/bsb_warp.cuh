// bsb_warp.cuh -- warp-cooperative kernels bodies (device only).
//
// The v1 kernels ran one read per thread; ncu showed 4-7 of 32 lanes active per issued instruction
// (profiles/r01_ncu_v1_thread_per_read.md). Here one WARP owns one read: the control logic is executed
// uniformly by all 32 lanes (no divergence), the dynamic programming rows are spread across the lanes.
//
// sw_extend_warp() is ksw_extend2 (ksw.c:380-479) row by row, exactly: the (h,e) row persists in shared
// memory with its stale cells, the band [beg,end) evolves per row from the zero cells, the row maximum takes
// the LAST column holding it, z-drop and the to-end score are evaluated per row. Inside a row
//     M(j)   = H(i-1,j-1) ? H(i-1,j-1) + s(i,j) : 0           (previous row only)
//     E(i,j)                                                   (previous row only)
//     F(i,j) = max_{j'<j} ( max(M(j') - oe_ins, 0) - (j-1-j') * e_ins )      F(i,beg) = 0
// so F is an exclusive max-scan of g(j') = max(M(j')-oe_ins,0) + j'*e_ins: five shuffle steps per 32 cells.
#pragma once
#include "bsb_stages.h"

namespace bsb {

#define FULLMASK 0xffffffffu
#define NEG_BIG (-0x3fffffff)

struct WarpDp {            // per-warp shared memory
    int32_t *H, *E;        // max_q + 1 each
    uint8_t *qs;           // query of the current extension, in extension order
};

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(FULLMASK, v); }   // one REDUX.MAX

template <class T>
__device__ ExtResult sw_extend_warp(int qlen, const QrySeq &query, int tlen, const T &target, const int8_t *mat,
                                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                                    const WarpDp &S)
{
    const int lane = threadIdx.x & 31;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int32_t *H = S.H, *E = S.E;
    uint8_t *qs = S.qs;
    // first row + query staging
    const int H1 = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = lane; j <= qlen; j += 32) {
        int v = 0;
        if (j == 0) v = h0;
        else { v = H1 - (j - 1) * e_ins; if (v < 0) v = 0; }
        H[j] = v; E[j] = 0;
        if (j < qlen) qs[j] = (uint8_t)query(j);
    }
    int i, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off, beg, end;
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
    __syncwarp();
    int tcache = 0;                                 // target bases of rows [i & ~31, +32), one per lane
    for (i = 0; i < tlen; ++i) {
        if ((i & 31) == 0) tcache = i + lane < tlen ? target(i + lane) : 4;
        const int8_t *row = mat + __shfl_sync(FULLMASK, tcache, i & 31) * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        int carry_g = NEG_BIG, carry_h = h1;       // running F source, H(i, j-1) entering the chunk
        int best = -1;                              // (h << 12 | j) maximum of the row
        int first_nz = -1, last_nz = -1;            // non-zero cells of the updated row in [beg,end)
        int phi = 0;                                // what any later row can still reach (ext_rows_exhausted, bsb_ksw.h)
        const bool at_qend = end == qlen;
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int j = c0 + lane;
            const bool act = j < end;
            int Hj = 0, Ej = 0, M = 0;
            if (act) { Hj = H[j]; Ej = E[j]; M = Hj ? Hj + row[qs[j]] : 0; }
            int tI = M - oe_ins; tI = tI > 0 ? tI : 0;
            int g = act ? tI + j * e_ins : NEG_BIG;
            int incl = g;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int v = __shfl_up_sync(FULLMASK, incl, d);
                if (lane >= d) incl = incl > v ? incl : v;
            }
            int excl = __shfl_up_sync(FULLMASK, incl, 1);
            if (lane == 0) excl = NEG_BIG;
            excl = excl > carry_g ? excl : carry_g;
            int f = (j == beg) ? 0 : excl - (j - 1) * e_ins;
            int h = M > Ej ? M : Ej;
            h = h > f ? h : f;
            int tD = M - oe_del; tD = tD > 0 ? tD : 0;
            int e2 = Ej - e_del; e2 = e2 > tD ? e2 : tD;
            int hprev = __shfl_up_sync(FULLMASK, h, 1);
            if (lane == 0) hprev = carry_h;
            if (act) { H[j] = hprev; E[j] = e2; }
            if (at_qend && act) {                   // per lane; reduced once per row below
                int p1 = hprev > 0 ? hprev + amax * (qlen - j) : 0, p2 = e2 > 0 ? e2 + amax * (qlen - 1 - j) : 0;
                p1 = p1 > p2 ? p1 : p2;
                phi = phi > p1 ? phi : p1;
            }
            int key = act ? (h << 12 | j) : -1;
            key = warp_max(key);
            best = best > key ? best : key;
            unsigned nz = __ballot_sync(FULLMASK, act && (hprev != 0 || e2 != 0));
            if (nz) {
                if (first_nz < 0) first_nz = c0 + __ffs(nz) - 1;
                last_nz = c0 + 31 - __clz(nz);
            }
            const int n_act = end - c0 < 32 ? end - c0 : 32;
            const int chunk_max = __shfl_sync(FULLMASK, incl, 31); // inactive lanes hold NEG_BIG
            carry_g = carry_g > chunk_max ? carry_g : chunk_max;
            carry_h = __shfl_sync(FULLMASK, h, n_act - 1);
        }
        h1 = carry_h;
        if (lane == 0) { H[end] = h1; E[end] = 0; }
        int m = 0, mj = -1;
        if (best >= 0) { m = best >> 12; mj = best & 0xfff; }
        if (end == qlen) {
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
        if (at_qend && ext_rows_exhausted(warp_max(phi), max, gscore)) break;
        // next band: drop leading/trailing cells whose h and e are both zero
        int nb = first_nz >= 0 ? first_nz : end;
        int jl;
        if (h1 != 0) jl = end;
        else if (last_nz >= 0 && last_nz >= nb) jl = last_nz;
        else jl = nb - 1;
        beg = nb;
        end = jl + 2 < qlen ? jl + 2 : qlen;
        __syncwarp();
    }
    ExtResult r;
    r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
    return r;
}

// ---------------------------------------------------------------------------------------------
// K4, one warp per read. Chains live in the lanes (two per lane, so up to 64 per read):
//   * the B-tree of the reference (kbtree.h, restated in bsb_chain.h) is only a map from a chain's first reference
//     position to the chain; with DISTINCT keys "closest chain at or before x" is a predecessor query -- here one
//     compare per lane and a warp maximum -- and the in-order traversal is the ascending order of the keys. A
//     duplicate key (a seed that starts exactly where an existing chain starts and cannot be merged into it), whose
//     place depends on the shape of the tree, sends the read to the serial form (stage_chain) instead;
//   * test_and_merge (bwamem.c:194-215) runs on the lane that owns the chain; the chain weight (mem_chain_weight,
//     bwamem.c:217-236) is accumulated as seeds are appended, in the same seed order;
//   * mem_chain_flt (bwamem.c:331-389) runs on lane 0 over 24-byte records in shared memory -- the introsort is the
//     order-exact restatement of bsb_hd.h, so the permutation among equal weights is the reference's.
// ---------------------------------------------------------------------------------------------
struct ChainRec { int32_t w, ci, beg, end, first; int8_t kept, is_alt; int16_t pad_; };
struct LtChainRecW { __device__ bool operator()(const ChainRec &a, const ChainRec &b) const { return a.w > b.w; } };
constexpr int CHAIN_SLOTS = 2, CHAIN_MAX = 32 * CHAIN_SLOTS;

struct LaneChain {
    int64_t pos, l_rbeg, endr;
    int32_t rid, n, head, tail, f_qbeg, l_qbeg, l_len, wq, endq, wr;
    int8_t is_alt; bool valid;
};

// maximum of non-negative 64-bit values (or -1 for "none") across the warp: two 32-bit REDUX steps
__device__ __forceinline__ int64_t warp_max_i64(int64_t v)
{
    const uint64_t u = (uint64_t)(v + 1);                       // 0 = none
    const uint32_t hi = __reduce_max_sync(FULLMASK, (uint32_t)(u >> 32));
    const uint32_t lo = __reduce_max_sync(FULLMASK, (uint32_t)(u >> 32) == hi ? (uint32_t)u : 0u);
    return (int64_t)((uint64_t)hi << 32 | lo) - 1;
}
struct LtKeyHiDesc { __device__ bool operator()(uint64_t a, uint64_t b) const { return (int32_t)(a >> 32) > (int32_t)(b >> 32); } };
constexpr int CHAIN_STAGE = 64;   // seeds staged in shared memory at a time

// test_and_merge (bwamem.c:194-215) on the lane that owns chain k; true when the seed was absorbed or appended
__device__ __forceinline__ bool lane_chain_merge(const Opt &opt, int64_t l_pac, LaneChain &k, const Seed &s, int si, int32_t *next)
{
    const int64_t qend = k.l_qbeg + k.l_len, rend = k.l_rbeg + k.l_len;
    if (s.rid != k.rid) return false;
    if (s.qbeg >= k.f_qbeg && s.qbeg + s.len <= qend && s.rbeg >= k.pos && s.rbeg + s.len <= rend) return true;   // contained
    if ((k.l_rbeg < l_pac || k.pos < l_pac) && s.rbeg >= l_pac) return false;                                     // other strand
    const int64_t x = s.qbeg - k.l_qbeg, y = s.rbeg - k.l_rbeg;
    if (!(y >= 0 && x - y <= opt.w && y - x <= opt.w && x - k.l_len < opt.max_chain_gap && y - k.l_len < opt.max_chain_gap)) return false;
    next[k.tail] = si; next[si] = -1;
    k.tail = si; ++k.n;
    k.l_qbeg = s.qbeg; k.l_rbeg = s.rbeg; k.l_len = s.len;
    // mem_chain_weight (bwamem.c:217-236), one seed at a time
    if (s.qbeg >= k.endq) k.wq += s.len; else if (s.qbeg + s.len > k.endq) k.wq += s.qbeg + s.len - k.endq;
    k.endq = k.endq > s.qbeg + s.len ? k.endq : s.qbeg + s.len;
    if (s.rbeg >= k.endr) k.wr += s.len; else if (s.rbeg + s.len > k.endr) k.wr += (int)(s.rbeg + s.len - k.endr);
    k.endr = k.endr > s.rbeg + s.len ? k.endr : s.rbeg + s.len;
    return true;
}

__device__ __forceinline__ void lane_chain_start(const IndexView &ix, LaneChain &k, const Seed &s, int si, int32_t *next)
{
    k.valid = true; k.pos = s.rbeg; k.rid = s.rid; k.n = 1; k.head = k.tail = si;
    k.f_qbeg = k.l_qbeg = s.qbeg; k.l_rbeg = s.rbeg; k.l_len = s.len;
    k.wq = s.len; k.endq = s.qbeg + s.len; k.wr = s.len; k.endr = s.rbeg + s.len;
    k.is_alt = (int8_t)(ix.anns[s.rid].is_alt != 0);
    next[si] = -1;
}

__device__ __forceinline__ int lane_chain_weight(const LaneChain &k)
{
    const int v = k.wq < k.wr ? k.wq : k.wr;
    return v < (1 << 30) ? v : (1 << 30) - 1;
}

__device__ __forceinline__ void lane_chain_emit(const BatchDev &B, uint32_t so, const Seed *seeds, const int32_t *next, const LaneChain &k,
                                                const ChainRec &t, int head, int out_idx, float frac_rep)
{
    Seed *cs = B.cseeds + so;
    int o = head;
    for (int j = k.head; j >= 0; j = next[j]) cs[o++] = seeds[j];
    Chain out;
    out.pos = k.pos; out.n = k.n; out.head = head; out.tail = o - 1; out.rid = k.rid; out.first = t.first;
    out.w = t.w; out.kept = t.kept; out.is_alt = k.is_alt; out.frac_rep = frac_rep;
    B.chains[so + out_idx] = out;
}

// K4 in three phases over a GROUP of 32 consecutive reads, so that the inherently serial parts run on 32 lanes at once
// instead of on lane 0 of a warp that waits:
//   phase 1 (warp per read, one read after the other): the seed loop above; every chain leaves a preliminary record in
//            chain_pool[so + rank] (rank = order of the first positions = in-order traversal of the reference's tree),
//            its query span in tmp[] and a (weight, rank) sort key in srt[];
//   phase 2 (lane per read): the order-exact introsort of the keys and mem_chain_flt over the records, through the
//            permutation held in the low halves of the keys; the surviving ranks are left in srt[0 .. n_out). Reads the
//            warp form cannot take (duplicate tree key, more than CHAIN_MAX chains) run the serial form on their lane here;
//   phase 3 (warp per read): offsets of the surviving chains in cseeds by a warp scan, every chain written out by one lane.

// phase 1: returns the number of chains that entered the sort (>= 0) or -1 when the read needs the serial form
__device__ int chain_phase1_warp(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, Seed *stage)
{
    const int lane = threadIdx.x & 31;
    const uint32_t so = B.seed_off[r];
    const int ns = (int)(B.seed_off[r + 1] - so);
    const Seed *seeds = B.seeds + so;
    int32_t *next = B.next + so;
    LaneChain c0, c1;          // chain `ci` lives in lane ci & 31, in c0 when ci < 32 else c1 (two named structs: registers)
    c0.valid = c1.valid = false;
    int n_chains = 0;
    const int64_t l_pac = ix.l_pac;
    for (int si = 0; si < ns; ++si) {
        if (si % CHAIN_STAGE == 0) {          // stage the next seeds (coalesced), then every lane reads them from shared memory
            __syncwarp();
            const int m = ns - si < CHAIN_STAGE ? ns - si : CHAIN_STAGE;
            const uint2 *src = reinterpret_cast<const uint2 *>(seeds + si);
            uint2 *dst = reinterpret_cast<uint2 *>(stage);
            for (int t = lane; t < m * 3; t += 32) dst[t] = src[t];
            __syncwarp();
        }
        const Seed s = stage[si % CHAIN_STAGE];
        if (s.rid < 0) continue;
        // predecessor query: the chain with the largest first position <= s.rbeg
        int64_t best = -1;
        if (c0.valid && c0.pos <= s.rbeg) best = c0.pos;
        if (c1.valid && c1.pos <= s.rbeg && c1.pos > best) best = c1.pos;
        best = warp_max_i64(best);
        bool merged = false;
        if (best >= 0) {
            if (c0.valid && c0.pos == best) merged = lane_chain_merge(opt, l_pac, c0, s, si, next);
            else if (c1.valid && c1.pos == best) merged = lane_chain_merge(opt, l_pac, c1, s, si, next);
        }
        if (__any_sync(FULLMASK, merged)) continue;
        if (best == s.rbeg || n_chains >= CHAIN_MAX) return -1;        // duplicate key / too many chains: serial form
        const int ci = n_chains++;
        if (lane == (ci & 31)) {
            if (ci < 32) lane_chain_start(ix, c0, s, si, next);
            else lane_chain_start(ix, c1, s, si, next);
        }
    }
    // tree order = ascending first position; chains lighter than min_chain_weight are dropped before the sort
    const int w0 = c0.valid ? lane_chain_weight(c0) : 0, w1 = c1.valid ? lane_chain_weight(c1) : 0;
    const bool keep0 = c0.valid && w0 >= opt.min_chain_weight, keep1 = c1.valid && w1 >= opt.min_chain_weight;
    int rank0 = 0, rank1 = 0, n_kept = 0;
    for (int ci = 0; ci < n_chains; ++ci) {
        int64_t p = ci < 32 ? c0.pos : c1.pos;
        int kp = ci < 32 ? (int)keep0 : (int)keep1;
        p = __shfl_sync(FULLMASK, p, ci & 31);
        kp = __shfl_sync(FULLMASK, kp, ci & 31);
        if (!kp) continue;
        ++n_kept;
        rank0 += p < c0.pos; rank1 += p < c1.pos;
    }
    Chain *pre = B.chain_pool + so;
    if (keep0) {
        Chain t; t.pos = c0.pos; t.n = c0.n; t.head = c0.head; t.tail = c0.tail; t.rid = c0.rid; t.first = -1; t.w = w0; t.kept = 0; t.is_alt = c0.is_alt; t.frac_rep = 0.f;
        pre[rank0] = t;
        B.tmp[so + rank0] = c0.f_qbeg | (c0.l_qbeg + c0.l_len) << 16;
        B.srt[so + rank0] = (uint64_t)(uint32_t)w0 << 32 | (uint32_t)rank0;
    }
    if (keep1) {
        Chain t; t.pos = c1.pos; t.n = c1.n; t.head = c1.head; t.tail = c1.tail; t.rid = c1.rid; t.first = -1; t.w = w1; t.kept = 0; t.is_alt = c1.is_alt; t.frac_rep = 0.f;
        pre[rank1] = t;
        B.tmp[so + rank1] = c1.f_qbeg | (c1.l_qbeg + c1.l_len) << 16;
        B.srt[so + rank1] = (uint64_t)(uint32_t)w1 << 32 | (uint32_t)rank1;
    }
    __syncwarp();
    return n_kept;
}

// phase 2, one lane per read: introsort by weight + mem_chain_flt (bwamem.c:331-389). a[i] of the reference is
// pre[(uint32_t)keys[i]]; keptl = this read's slice of `aux`. Returns the number of chains that survive.
__device__ int chain_phase2_lane(const Opt &opt, const BatchDev &B, int r, int n_chn, int32_t *aux)
{
    const uint32_t so = B.seed_off[r];
    uint64_t *keys = B.srt + so;
    Chain *pre = B.chain_pool + so;
    const int32_t *span = B.tmp + so;
    int32_t *keptl = aux + so;
    if (n_chn <= 0) return 0;
    introsort((long)n_chn, keys, LtKeyHiDesc());
#define BSB_A(i) pre[(uint32_t)keys[i]]
    int nk = 0, i, k;
    BSB_A(0).kept = 3;
    keptl[nk++] = 0;
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        const uint32_t pi = (uint32_t)keys[i];
        const int beg_i = span[pi] & 0xffff, end_i = span[pi] >> 16, w_i = pre[pi].w, alt_i = pre[pi].is_alt;
        for (k = 0; k < nk; ++k) {
            const int j = keptl[k];
            const uint32_t pj = (uint32_t)keys[j];
            const int beg_j = span[pj] & 0xffff, end_j = span[pj] >> 16;
            const int b_max = beg_j > beg_i ? beg_j : beg_i;
            const int e_min = end_j < end_i ? end_j : end_i;
            if (e_min > b_max && (!pre[pj].is_alt || alt_i)) {
                const int li = end_i - beg_i, lj = end_j - beg_j;
                const int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
                    large_ovlp = 1;
                    if (pre[pj].first < 0) pre[pj].first = i;
                    const int w_j = pre[pj].w;
                    if (w_i < w_j * opt.drop_ratio && w_j - w_i >= opt.min_seed_len << 1) break;
                }
            }
        }
        if (k == nk) {
            keptl[nk++] = i;
            pre[pi].kept = large_ovlp ? 2 : 3;
        }
    }
    for (i = 0; i < nk; ++i) {
        const Chain &t = BSB_A(keptl[i]);
        if (t.first >= 0) BSB_A(t.first).kept = 1;
    }
    for (i = k = 0; i < n_chn; ++i) {
        if (BSB_A(i).kept == 0 || BSB_A(i).kept == 3) continue;
        if (++k >= opt.max_chain_extend) break;
    }
    for (; i < n_chn; ++i)
        if (BSB_A(i).kept < 3) BSB_A(i).kept = 0;
    for (i = k = 0; i < n_chn; ++i)
        if (BSB_A(i).kept != 0) keys[k++] = (uint32_t)keys[i];       // k <= i: the surviving ranks, in sorted order
#undef BSB_A
    return k;
}

// phase 3: the surviving chains of read r (ranks in srt[0 .. n_out)) go to chains[] / cseeds[], one chain per lane
__device__ void chain_phase3_warp(const BatchDev &B, int r, int n_out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t so = B.seed_off[r];
    const uint64_t *keys = B.srt + so;
    const Chain *pre = B.chain_pool + so;
    const Seed *seeds = B.seeds + so;
    const int32_t *next = B.next + so;
    const float frac_rep = (float)B.l_rep[r] / (float)(int)(B.seq_off[r + 1] - B.seq_off[r]);
    int run = 0;
    for (int k0 = 0; k0 < n_out; k0 += 32) {
        const int k = k0 + lane;
        Chain t;
        t.n = 0;
        if (k < n_out) t = pre[(uint32_t)keys[k]];
        int incl = t.n;                                    // offsets: exclusive sum of the seed counts, in output order
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += v; }
        const int head = run + incl - t.n;
        run += __shfl_sync(FULLMASK, incl, 31);
        if (k < n_out) {
            Seed *cs = B.cseeds + so;
            int o = head;
            for (int j = t.head; j >= 0; j = next[j]) cs[o++] = seeds[j];
            Chain out = t;
            out.head = head; out.tail = o - 1; out.frac_rep = frac_rep;
            B.chains[so + k] = out;
        }
    }
    if (lane == 0) B.n_chain[r] = n_out;
}

struct ChainSmem { Seed stage[CHAIN_STAGE]; };

// one group of up to 32 consecutive reads [base, base + 32)
__device__ void stage_chain_group(const Opt &opt, const IndexView &ix, const BatchDev &B, int base, ChainSmem &S, int32_t *aux)
{
    const int lane = threadIdx.x & 31;
    const int cnt = B.n - base < 32 ? B.n - base : 32;
    int my_n = 0;                                          // lane q: chains of read base + q that enter the sort; -1: nothing to do
    for (int q = 0; q < cnt; ++q) {
        const int r = base + q;
        const int ns = (int)(B.seed_off[r + 1] - B.seed_off[r]);
        int n = -1;
        if (ns == 0 || B.err[r]) { if (lane == 0) B.n_chain[r] = 0; }
        else {
            n = chain_phase1_warp(opt, ix, B, r, S.stage);
            if (n < 0) n = -2;                             // the serial form, by this read's lane in phase 2
            else if (n == 0) { if (lane == 0) B.n_chain[r] = 0; n = -1; }
        }
        if (lane == q) my_n = n;
    }
    __syncwarp();
    int my_out = 0;
    if (lane < cnt && my_n > 0) my_out = chain_phase2_lane(opt, B, base + lane, my_n, aux);
    else if (lane < cnt && my_n == -2) stage_chain(opt, ix, B, base + lane);
    __syncwarp();
    for (int q = 0; q < cnt; ++q) {
        const int n = __shfl_sync(FULLMASK, my_n, q), n_out = __shfl_sync(FULLMASK, my_out, q);
        if (n > 0) chain_phase3_warp(B, base + q, n_out);
    }
}

// mem_chain2aln, warp-uniform: every lane executes the same control flow on the same values; lane 0 alone
// writes to HBM; the two extensions per seed run across the lanes.
__device__ void chain_to_regions_warp(const Opt &opt, const IndexView &ix, int l_query, const uint8_t *query,
                                      const Chain &c, const Seed *cs, uint64_t *srt, RegList &av, const WarpDp &S, DpScratch &dp, int *err)
{
    const int lane = threadIdx.x & 31;
    int i, k, max_off[2], aw[2];
    const int64_t l_pac = ix.l_pac;
    int64_t rmax[2], tmp;
    if (c.n == 0) return;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (i = 0; i < c.n; ++i) {
        const Seed &t = cs[i];
        int64_t b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
        int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
    if (rmax[0] < l_pac && l_pac < rmax[1]) {
        if (cs[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    fetch_window(ix, &rmax[0], cs[0].rbeg, &rmax[1]);
    if (l_query > dp.max_q) { *err = ERR_SCRATCH_OVERFLOW; return; }
    if (lane == 0) {
        for (i = 0; i < c.n; ++i) srt[i] = (uint64_t)cs[i].score << 32 | (uint32_t)i;
        introsort((long)c.n, srt, LtU64());
    }
    __syncwarp();
    for (k = c.n - 1; k >= 0; --k) {
        const Seed s = cs[(uint32_t)srt[k]];
        for (i = 0; i < av.n; ++i) {
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
            for (i = k + 1; i < c.n; ++i) {
                if (srt[i] == 0) continue;
                const Seed &t = cs[(uint32_t)srt[i]];
                if (t.len < s.len * .95) continue;
                if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
            }
            if (i == c.n) {
                __syncwarp();
                if (lane == 0) srt[k] = 0;
                __syncwarp();
                continue;
            }
        }
        if (av.n >= av.cap) { *err = ERR_SCRATCH_OVERFLOW; return; }
        AlnReg a;
        alnreg_clear(a);
        a.w = aw[0] = aw[1] = opt.w;
        a.score = a.truesc = -1;
        a.rid = c.rid;
        if (s.qbeg) {
            QrySeq qs = {query + (s.qbeg - 1), -1};
            RefSeq rs = {ix.pac, l_pac, s.rbeg - 1, -1};
            tmp = s.rbeg - rmax[0];
            ExtResult x = {0, 0, 0, 0, 0, 0};
            for (i = 0; i < 2; ++i) {
                int prev = a.score;
                aw[0] = opt.w << i;
                x = sw_extend_warp(s.qbeg, qs, (int)tmp, rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[0], opt.pen_clip5, opt.zdrop, s.len * opt.a, S);
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
        if (s.qbeg + s.len != l_query) {
            int qe = s.qbeg + s.len, sc0 = a.score;
            int64_t re = s.rbeg + s.len - rmax[0];
            QrySeq qs = {query + qe, 1};
            RefSeq rs = {ix.pac, l_pac, rmax[0] + re, 1};
            ExtResult x = {0, 0, 0, 0, 0, 0};
            for (i = 0; i < 2; ++i) {
                int prev = a.score;
                aw[1] = opt.w << i;
                x = sw_extend_warp(l_query - qe, qs, (int)(rmax[1] - rmax[0] - re), rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[1], opt.pen_clip3, opt.zdrop, sc0, S);
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
        __syncwarp();
        if (lane == 0) av.a[av.n] = a;
        ++av.n;
        __syncwarp();
    }
}

// K5. A warp takes a group of up to 32 reads: the chains of each read are extended by the whole warp, one read after the
// other (phase A); the tail of mem_align1_core -- mem_sort_dedup_patch with its introsorts and the occasional score-only
// global alignment of mem_patch_reg, serial by nature -- then runs for all reads of the group at once, one read per lane,
// each lane with its own DP scratch (phase B), instead of on lane 0 of a waiting warp after every read.
__device__ void stage_extend_group(const Opt &opt, const IndexView &ix, const BatchDev &B, int base, int cnt, const int32_t *order,
                                   const WarpDp &S, DpScratch &dp_warp, DpScratch &dp_lane)
{
    const int lane = threadIdx.x & 31;
    int my_n = -1, my_err = 0;                             // lane q: regions of read base + q before the tail (-1: nothing to do)
    for (int q = 0; q < cnt; ++q) {
        const int r = order ? order[base + q] : base + q;
        const uint32_t so = B.seed_off[r];
        const int ns = (int)(B.seed_off[r + 1] - so);
        int n = -1, err = 0;
        if (ns == 0 || B.err[r]) { if (lane == 0) B.n_regs[r] = 0; }
        else {
            const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
            const uint8_t *seq = B.seq + B.seq_off[r];
            RegList av = {B.regs + so, 0, ns};
            const int nc = B.n_chain[r];
            for (int i = 0; i < nc; ++i) {
                const Chain c = B.chains[so + i];
                chain_to_regions_warp(opt, ix, len, seq, c, B.cseeds + so + c.head, B.srt + so, av, S, dp_warp, &err);
                if (err) break;
            }
            n = av.n;
        }
        if (lane == q) { my_n = n; my_err = err; }
    }
    __syncwarp();
    if (lane < cnt && my_n >= 0) {
        const int r = order ? order[base + lane] : base + lane;
        AlnReg *a = B.regs + B.seed_off[r];
        int n = my_n, err = my_err;
        if (!err) {
            n = sort_dedup_patch(opt, ix, B.seq + B.seq_off[r], n, a, dp_lane, &err);
            for (int i = 0; i < n; ++i)
                if (a[i].rid >= 0 && ix.anns[a[i].rid].is_alt) a[i].is_alt = 1;
        }
        if (err) { B.err[r] = err; B.n_regs[r] = 0; }
        else B.n_regs[r] = n;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// K6: banded global alignment + traceback, one warp per queued alignment (AlnTask).
//
// nw_global_warp() is ksw_global2 (ksw.c:504-606) row by row with the columns across the lanes: M and E
// come from the previous row, F(i,j) = max( -inf - (j-beg)*e_ins, max_{j'<j} (M(j') - oe_ins - (j-1-j')*e_ins) )
// is an exclusive max-scan, and each lane derives its own three direction bits with the reference's tie
// rules (m >= e, h >= f, e > t, f > t). The direction matrix goes to HBM (coalesced rows), the backtrack
// walks it on lane 0.
// ---------------------------------------------------------------------------------------------
struct WarpTask {          // per-warp shared memory of the task kernel
    int32_t *H, *E;        // max_q + 1 each
    uint8_t *qs;           // max_q
    uint32_t *cigar; int cigar_cap;
    char *md; int md_cap;
    char *xb; int xb_cap;
};

template <class Q, class T>
__device__ int nw_global_warp(int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                              int o_del, int e_del, int o_ins, int e_ins, int w, const WarpTask &S, uint8_t *z, CigarBuf *cig, int *err)
{
    const int lane = threadIdx.x & 31;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int32_t *H = S.H, *E = S.E;
    uint8_t *qs = S.qs;
    for (int j = lane; j <= qlen; j += 32) {
        int h;
        if (j == 0) h = 0;
        else if (j <= w) h = -(o_ins + e_ins * j);
        else h = BSB_MINUS_INF;
        H[j] = h; E[j] = BSB_MINUS_INF;
        if (j < qlen) qs[j] = (uint8_t)query(j);
    }
    __syncwarp();
    int tcache = 0;                                 // target bases of rows [i & ~31, +32), one per lane
    for (int i = 0; i < tlen; ++i) {
        if ((i & 31) == 0) tcache = i + lane < tlen ? target(i + lane) : 4;
        const int8_t *row = mat + __shfl_sync(FULLMASK, tcache, i & 31) * 5;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        int carry_h = beg == 0 ? -(o_del + e_del * (i + 1)) : BSB_MINUS_INF;
        int carry_g = BSB_MINUS_INF + (beg - 1) * e_ins;   // the "-inf" F entering column beg, in scan coordinates
        uint8_t *zi = z + (long)i * n_col;
        for (int c0 = beg; c0 < end; c0 += 32) {
            const int j = c0 + lane;
            const bool act = j < end;
            int m = 0, e = 0;
            if (act) { m = H[j] + row[qs[j]]; e = E[j]; }
            int g = act ? m - oe_ins + j * e_ins : BSB_MINUS_INF * 2 + 1;
            int incl = g;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int v = __shfl_up_sync(FULLMASK, incl, d);
                if (lane >= d) incl = incl > v ? incl : v;
            }
            int excl = __shfl_up_sync(FULLMASK, incl, 1);
            if (lane == 0) excl = carry_g;
            else excl = excl > carry_g ? excl : carry_g;
            const int f = excl - (j - 1) * e_ins;
            uint8_t d = m >= e ? 0 : 1;
            int h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            int t = m - oe_del;
            int e2 = e - e_del;
            d |= e2 > t ? 1 << 2 : 0;
            e2 = e2 > t ? e2 : t;
            t = m - oe_ins;
            const int f2 = f - e_ins;
            d |= f2 > t ? 2 << 4 : 0;
            int hprev = __shfl_up_sync(FULLMASK, h, 1);
            if (lane == 0) hprev = carry_h;
            if (act) { H[j] = hprev; E[j] = e2; zi[j - beg] = d; }
            const int n_act = end - c0 < 32 ? end - c0 : 32;
            const int chunk_max = __shfl_sync(FULLMASK, incl, n_act - 1);
            carry_g = carry_g > chunk_max ? carry_g : chunk_max;
            carry_h = __shfl_sync(FULLMASK, h, n_act - 1);
        }
        if (lane == 0) { H[end] = carry_h; E[end] = BSB_MINUS_INF; }
        __syncwarp();
    }
    const int score = H[qlen];
    int n_cig = 0, bad = 0;
    if (lane == 0) {
        cig->n = 0;
        int which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        bool ok = true;
        while (i >= 0 && k >= 0) {
            which = z[(long)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { ok &= cig->push(0, 1); --i; --k; }
            else if (which == 1) { ok &= cig->push(2, 1); --i; }
            else { ok &= cig->push(1, 1); --k; }
        }
        if (i >= 0) ok &= cig->push(2, i + 1);
        if (k >= 0) ok &= cig->push(1, k + 1);
        for (i = 0; i < cig->n >> 1; ++i) tswap(cig->a[i], cig->a[cig->n - 1 - i]);
        n_cig = cig->n; bad = ok ? 0 : 1;
    }
    n_cig = __shfl_sync(FULLMASK, n_cig, 0);
    bad = __shfl_sync(FULLMASK, bad, 0);
    cig->n = n_cig;
    if (bad) *err = ERR_CIGAR_OVERFLOW;
    __syncwarp();
    return score;
}

// bwa_gen_cigar2 part 1 (bwa.c:199-249), warp form; all lanes pass identical arguments
__device__ bool global_core_warp(const Opt &opt, const IndexView &ix, int w_, int l_query, const uint8_t *query,
                                 int64_t rb, int64_t re, int *score, CigarBuf *cig, const WarpTask &S, uint8_t *z, long z_cap, int max_q, int *err)
{
    const int lane = threadIdx.x & 31;
    const int64_t l_pac = ix.l_pac;
    cig->n = 0;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
    if (re > (l_pac << 1) || rb < 0) return false;
    const int64_t rlen = re - rb;
    QrySeq q; RefSeq t;
    if (rb >= l_pac) {
        q.base = query + (l_query - 1); q.dir = -1;
        t.pac = ix.pac; t.l_pac = l_pac; t.start = re - 1; t.dir = -1;
    } else {
        q.base = query; q.dir = 1;
        t.pac = ix.pac; t.l_pac = l_pac; t.start = rb; t.dir = 1;
    }
    if (l_query == rlen && w_ == 0) {
        if (lane == 0) { cig->a[0] = (uint32_t)l_query << 4; }
        cig->n = 1;
        int sc = 0;
        for (int i = lane; i < l_query; i += 32) sc += opt.mat[t(i) * 5 + q(i)];
        for (int o = 16; o; o >>= 1) sc += __shfl_xor_sync(FULLMASK, sc, o);
        *score = sc;
        __syncwarp();
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
        if (l_query > max_q) { *err = ERR_SCRATCH_OVERFLOW; return false; }
        long n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
        if (n_col * rlen > z_cap) { *err = ERR_SCRATCH_OVERFLOW; return false; }
        *score = nw_global_warp(l_query, q, (int)rlen, t, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, S, z, cig, err);
    }
    return true;
}

// A queued alignment needs no dynamic programming when the band inferred for it is empty and both spans have the
// same length: bwa_gen_cigar2 then returns <len>M whatever the score (bwa.c:222-230), also on the retries.
__device__ __forceinline__ bool task_is_trivial(const Opt &opt, const AlnTask &t)
{
    return band_for_task(opt, t) == 0 && (int64_t)(t.qe - t.qb) == t.re - t.rb;
}

// K6a: the global alignments. Each warp scans 32 queued alignments at a time and runs the ones that need the
// DP, one after the other, with its 32 lanes across the columns; the CIGAR goes to the task's slot in HBM.
__device__ void stage_task_dp_warp(const Opt &opt, const IndexView &ix, BatchDev &B, unsigned int k, const WarpTask &S, uint8_t *z, long z_cap, int max_q,
                                   uint32_t *slot, int slot_cap, int32_t *n_cig_out)
{
    const int lane = threadIdx.x & 31;
    const AlnTask t = B.tasks.a[k];
    const int r = t.read;
    const uint8_t *query = B.seq + B.seq_off[r];
    int err = 0;
    int w2 = band_for_task(opt, t), score = 0, last_sc = -(1 << 30), i = 0;
    CigarBuf cig = {slot, 0, slot_cap - 2};
    bool ok;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        ok = global_core_warp(opt, ix, w2, t.qe - t.qb, query + t.qb, t.rb, t.re, &score, &cig, S, z, z_cap, max_q, &err);
        if (!ok) break;
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < t.truesc - opt.a);
    if (lane == 0) {
        *n_cig_out = ok ? cig.n : -1;
        if (err) B.out[r].err = err;
    }
    __syncwarp();
}

// K6b + K8b: one queued alignment per THREAD once its CIGAR exists: bisulfite NM/MD/XB against the unconverted
// reference, position, clips, and the write into the record(s) in the arena.
__device__ void stage_task_finish(const Opt &opt, const IndexView &ix, BatchDev &B, unsigned int k, uint32_t *slot, const int32_t *n_cig_in,
                                  char *md_buf, int md_cap, char *xb_buf, int xb_cap)
{
    const AlnTask t = B.tasks.a[k];
    const int r = t.read;
    if (B.out[r].err) return;
    const int l = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    const uint8_t *oquery = B.oseq + B.seq_off[r];
    uint32_t small[4];
    uint32_t *c = slot;
    int n_cig, err = 0;
    if (task_is_trivial(opt, t)) {
        const int l_query = t.qe - t.qb;
        const bool valid = !(l_query <= 0 || t.rb >= t.re || (t.rb < ix.l_pac && t.re > ix.l_pac)) && !(t.re > (ix.l_pac << 1) || t.rb < 0);
        c = small; c[0] = (uint32_t)l_query << 4; n_cig = valid ? 1 : -1;
    } else n_cig = *n_cig_in;
    if (n_cig < 0) err = ERR_NO_MD;
    else {
        AlnBody b;
        StrBuf md = {md_buf, 0, md_cap, false}, xb = {xb_buf, 0, xb_cap, false};
        aln_finish(ix, t, l, oquery, c, n_cig, md, xb, b, &err);
        task_store(ix, t, b, c, md_buf, B.arena, B.out, &err);
    }
    if (err) B.out[r].err = err;
}

// One queued alignment on one warp: CIGAR by the lanes together, bisulfite MD/XB + record write by lane 0
__device__ void stage_task_warp(const Opt &opt, const IndexView &ix, BatchDev &B, unsigned int k, const WarpTask &S, uint8_t *z, long z_cap, int max_q)
{
    const int lane = threadIdx.x & 31;
    const AlnTask t = B.tasks.a[k];
    const int r = t.read;
    const int l = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    const uint8_t *query = B.seq + B.seq_off[r], *oquery = B.oseq + B.seq_off[r];
    int err = 0;
    int w2 = band_for_task(opt, t), score = 0, last_sc = -(1 << 30), i = 0;
    CigarBuf cig = {S.cigar, 0, S.cigar_cap - 2};
    bool ok;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        ok = global_core_warp(opt, ix, w2, t.qe - t.qb, query + t.qb, t.rb, t.re, &score, &cig, S, z, z_cap, max_q, &err);
        if (!ok) break;
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < t.truesc - opt.a);
    if (lane == 0) {
        if (!ok) { if (!err) err = ERR_NO_MD; }
        else {
            AlnBody b;
            StrBuf md = {S.md, 0, S.md_cap, false}, xb = {S.xb, 0, S.xb_cap, false};
            aln_finish(ix, t, l, oquery, S.cigar, cig.n, md, xb, b, &err);
            task_store(ix, t, b, S.cigar, S.md, B.arena, B.out, &err);
        }
        if (err) B.out[r].err = err;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// K7, heavy pairs: mate rescue with the Smith-Waterman spread over the warp.
//
// sw_striped_warp() is sw_striped() (bsb_ksw.h: the cell-exact restatement of ksw_u8 / ksw_i16, ksw.c:115-339) with the
// query positions of a row across the lanes. In the linear restatement the stripe-local F of the first sweep is a
// SEGMENTED exclusive max-scan (segments = stripes of slen cells) of g(q) = max(h(q) - oe_ins, 0) + q * e_ins, and the
// lazy-F repair is a plain exclusive max-scan of the same quantity over the row: a cell whose h was itself raised by F
// cannot raise a later F (oe_ins > e_ins), so both scans read the pre-F values. E and the row maximum use the
// first-sweep h, exactly like the reference.
// ---------------------------------------------------------------------------------------------
struct WarpSw {
    int32_t *H0, *H1, *E, *Hmax;   // shared memory, cap cells each
    uint8_t *qs;                   // shared memory, cap bytes: query codes of the current pass
    int cap;
    uint64_t *b; int cap_b;        // sub-optimal list (HBM)
};

template <class Q, class T>
__device__ SwResult sw_striped_warp(int size, int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                                    int o_del, int e_del, int o_ins, int e_ins, int xtra, const WarpSw &ws, int *err)
{
    const int lane = threadIdx.x & 31;
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
    const int minsc = (xtra & SW_XSUBO) ? xtra & 0xffff : 0x10000;
    const int endsc = (xtra & SW_XSTOP) ? xtra & 0xffff : 0x10000;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int32_t *H0 = ws.H0, *H1 = ws.H1, *E = ws.E, *Hmax = ws.Hmax;
    uint8_t *qs = ws.qs;
    for (int q = lane; q < L; q += 32) { E[q] = H0[q] = Hmax[q] = 0; qs[q] = q < qlen ? (uint8_t)query(q) : (uint8_t)5; }
    int n_b = 0, te = -1, gmax = 0, last_b_sc = 0, last_b_i = -2, tcache = 0, i;
    __syncwarp();
    for (i = 0; i < tlen; ++i) {
        if ((i & 31) == 0) tcache = i + lane < tlen ? target(i + lane) : 4;
        const int8_t *row = mat + __shfl_sync(FULLMASK, tcache, i & 31) * 5;
        const uint64_t rowpack = (uint64_t)(uint8_t)row[0] | (uint64_t)(uint8_t)row[1] << 8 | (uint64_t)(uint8_t)row[2] << 16 |
                                 (uint64_t)(uint8_t)row[3] << 24 | (uint64_t)(uint8_t)row[4] << 32;   // byte 5 = 0: the pad cells
        int imax = 0;
        int carry1 = NEG_BIG, carry1_seg = -1;   // first sweep: best g of the stripe that runs into this chunk
        int carry2 = NEG_BIG;                    // lazy-F sweep: best g of everything before this chunk
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int q = c0 + lane;
            const bool act = q < L;
            int hh = 0, e = 0, seg = -2;
            if (act) {
                hh = q ? H0[q - 1] : 0;
                const int sc = (int)(int8_t)(rowpack >> (8 * qs[q]));
                if (size == 1) { hh = hh + sc + shift; if (hh > 255) hh = 255; hh = hh - shift; if (hh < 0) hh = 0; }
                else { hh = hh + sc; if (hh > 32767) hh = 32767; }
                e = E[q];
                hh = hh > e ? hh : e;
                seg = q / slen;
            }
            // first sweep: F restarts at the head of every stripe
            int t = hh - oe_ins; t = t > 0 ? t : 0;
            const int g = act ? t + q * e_ins : NEG_BIG;
            int incl = g;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULLMASK, incl, d);
                const int sg = __shfl_up_sync(FULLMASK, seg, d);
                if (lane >= d && sg == seg) incl = incl > v ? incl : v;
            }
            int excl = __shfl_up_sync(FULLMASK, incl, 1);
            const int pseg = __shfl_up_sync(FULLMASK, seg, 1);
            if (lane == 0 || pseg != seg) excl = NEG_BIG;
            if (seg == carry1_seg && carry1 > excl) excl = carry1;
            int f = excl - (q - 1) * e_ins; f = f > 0 ? f : 0;
            if (!act) f = 0;
            int h = hh > f ? hh : f;
            const int rm = __reduce_max_sync(FULLMASK, act ? h : 0);
            imax = imax > rm ? imax : rm;
            int e2 = e - e_del; e2 = e2 > 0 ? e2 : 0;
            int tD = h - oe_del; tD = tD > 0 ? tD : 0;
            if (act) E[q] = e2 > tD ? e2 : tD;
            // carry of the first sweep: the running maximum of the last lane's stripe
            const int last_seg = __shfl_sync(FULLMASK, seg, 31);
            const int last_incl = __shfl_sync(FULLMASK, incl, 31);
            int nc1 = last_incl;
            if (last_seg == carry1_seg && carry1 > nc1) nc1 = carry1;
            // does the last lane's stripe start inside this chunk? then the old carry does not belong to it (handled by the test above)
            carry1 = nc1; carry1_seg = last_seg;
            // lazy-F repair: plain scan over the row of the first-sweep h
            int t2 = h - oe_ins; t2 = t2 > 0 ? t2 : 0;
            const int g2 = act ? t2 + q * e_ins : NEG_BIG;
            int incl2 = g2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULLMASK, incl2, d);
                if (lane >= d) incl2 = incl2 > v ? incl2 : v;
            }
            int excl2 = __shfl_up_sync(FULLMASK, incl2, 1);
            if (lane == 0) excl2 = NEG_BIG;
            excl2 = excl2 > carry2 ? excl2 : carry2;
            int f2 = excl2 - (q - 1) * e_ins; f2 = f2 > 0 ? f2 : 0;
            if (act) H1[q] = f2 > h ? f2 : h;
            const int cm2 = __shfl_sync(FULLMASK, incl2, 31);
            carry2 = carry2 > cm2 ? carry2 : cm2;
        }
        __syncwarp();
        if (imax >= minsc) {
            if (n_b == 0 || last_b_i + 1 != i) {
                if (n_b >= ws.cap_b) { *err = ERR_SCRATCH_OVERFLOW; return r; }
                if (lane == 0) ws.b[n_b] = (uint64_t)imax << 32 | (uint32_t)i;
                ++n_b; last_b_sc = imax; last_b_i = i;
            } else if (last_b_sc < imax) {
                if (lane == 0) ws.b[n_b - 1] = (uint64_t)imax << 32 | (uint32_t)i;
                last_b_sc = imax; last_b_i = i;
            }
        }
        bool stop = false;
        if (imax > gmax) {
            gmax = imax; te = i;
            for (int q = lane; q < L; q += 32) Hmax[q] = H1[q];
            if (size == 1) { if (gmax + shift >= 255 || gmax >= endsc) stop = true; }
            else if (gmax >= endsc) stop = true;
        }
        if (stop) break;
        int32_t *S = H1; H1 = H0; H0 = S;
        __syncwarp();
    }
    __syncwarp();
    r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (size != 1 || r.score != 255) {
        // smallest query position holding the maximum of Hmax
        int key = -1;
        for (int q = lane; q < L; q += 32) { const int k = Hmax[q] << 12 | (4095 - q); key = key > k ? key : k; }
        key = __reduce_max_sync(FULLMASK, key);
        r.qe = key >= 0 ? 4095 - (key & 0xfff) : -1;
        int s2 = -1, te2 = -1;
        if (n_b && lane == 0) {
            int k = (r.score + qmax - 1) / qmax;
            const int low = te - k, high = te + k;
            for (k = 0; k < n_b; ++k) {
                const int e = (int32_t)ws.b[k];
                if ((e < low || e > high) && (int)(ws.b[k] >> 32) > s2) { s2 = (int)(ws.b[k] >> 32); te2 = e; }
            }
        }
        r.score2 = __shfl_sync(FULLMASK, s2, 0);
        r.te2 = __shfl_sync(FULLMASK, te2, 0);
    }
    __syncwarp();
    return r;
}

// ksw_align2 (ksw.c:343-365) on the warp: forward pass, then the reverse pass that locates the start
template <class Q, class T>
__device__ SwResult sw_local_warp(int qlen, const Q &query, int tlen, const T &target, const int8_t *mat,
                                  int o_del, int e_del, int o_ins, int e_ins, int xtra, const WarpSw &ws, int *err)
{
    const int size = (xtra & SW_XBYTE) ? 1 : 2;
    SwResult r = sw_striped_warp(size, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra, ws, err);
    if ((xtra & SW_XSTART) == 0 || ((xtra & SW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    struct RevQ { const Q &q; int n; __device__ int operator()(int i) const { return q(n - 1 - i); } };
    struct RevT { const T &t; int n; __device__ int operator()(int i) const { return i < n ? t(n - 1 - i) : t(i); } };
    RevQ rq = {query, r.qe + 1};
    RevT rt = {target, r.te + 1};
    SwResult rr = sw_striped_warp(size, r.qe + 1, rq, tlen, rt, mat, o_del, e_del, o_ins, e_ins, SW_XSTOP | r.score, ws, err);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}

// mem_matesw (bwamem_pair.c:111-180), warp-uniform: every lane follows the same control flow on the same values, the
// Smith-Waterman runs across the lanes, lane 0 alone modifies the mate's region list.
__device__ int mate_rescue_warp(const Opt &opt, const IndexView &ix, const PeStat pes[4], const AlnReg &a, int l_ms, const uint8_t *ms,
                                RegList &ma, FinalWS &ws, const WarpSw &sw, int *err)
{
    const int lane = threadIdx.x & 31;
    const int64_t l_pac = ix.l_pac;
    int i, r, skip[4], n = 0;
    for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
    for (i = 0; i < ma.n; ++i) {
        int64_t dist;
        r = infer_dir(l_pac, a.rb, ma.a[i].rb, &dist);
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
    }
    if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
    for (r = 0; r < 4; ++r) {
        int is_rev, is_larger, rid = -1;
        int64_t rb, re;
        if (skip[r]) continue;
        is_rev = (r >> 1 != (r & 1));
        is_larger = !(r >> 1);
        const uint8_t *seq = ms;
        if (is_rev) {
            __syncwarp();
            for (i = lane; i < l_ms; i += 32) ws.rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
            __syncwarp();
            seq = ws.rev;
        }
        if (!is_rev) {
            rb = is_larger ? a.rb + pes[r].low : a.rb - pes[r].high;
            re = (is_larger ? a.rb + pes[r].high : a.rb - pes[r].low) + l_ms;
        } else {
            rb = (is_larger ? a.rb + pes[r].low : a.rb - pes[r].high) - l_ms;
            re = is_larger ? a.rb + pes[r].high : a.rb - pes[r].low;
        }
        if (rb < 0) rb = 0;
        if (re > l_pac << 1) re = l_pac << 1;
        bool have_ref = false;
        if (rb < re) { rid = fetch_window(ix, &rb, (rb + re) >> 1, &re); have_ref = true; }
        if (have_ref && a.rid == rid && re - rb >= opt.min_seed_len) {
            int xtra = SW_XSUBO | SW_XSTART | (l_ms * opt.a < 250 ? SW_XBYTE : 0) | (opt.min_seed_len * opt.a);
            QrySeq q = {seq, 1};
            RefSeq t = {ix.pac, l_pac, rb, 1};
            SwResult aln = sw_local_warp(l_ms, q, (int)(re - rb), t, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, xtra, sw, err);
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
                __syncwarp();
                if (lane == 0) {
                    for (i = 0; i < ma.n - 1; ++i)
                        if (ma.a[i].score < b.score) break;
                    int tmp = i;
                    for (i = ma.n - 1; i > tmp; --i) ma.a[i] = ma.a[i - 1];
                    ma.a[i] = b;
                }
                __syncwarp();
            }
            ++n;
        }
        if (n) {
            int m = 0, e0 = 0;
            __syncwarp();
            if (lane == 0) { m = sort_dedup_patch(opt, ix, nullptr, ma.n, ma.a, ws.dp, &e0); }
            m = __shfl_sync(FULLMASK, m, 0); e0 = __shfl_sync(FULLMASK, e0, 0);
            ma.n = m;
            if (e0) *err = e0;
            __syncwarp();
        }
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// Replay of the rescue calls over precomputed Smith-Waterman results (bsb_final.h: RescueJob / RescuePre), one warp per
// pair with the pair's region lists in shared memory. What is left of mem_matesw once the alignments are known is
// mem_sort_dedup_patch after every call -- two sorts of a list that has a hundred entries for the pairs that get here. The
// reference's sorts are unstable introsorts, so the order of EQUAL keys is result-visible; with all keys distinct any sort
// gives the same array. The warp therefore ranks the entries by counting (every lane a few entries against all others),
// notices equal keys while it counts, and only then lets lane 0 run the order-exact introsort instead.
// ---------------------------------------------------------------------------------------------
template <class Less>
__device__ void warp_sort_regs(int n, AlnReg *a, AlnReg *tmp, Less lt)
{
    const int lane = threadIdx.x & 31;
    bool dup = false;
    for (int e = lane; e < n; e += 32) {
        const AlnReg x = a[e];
        int rank = 0;
        for (int f = 0; f < n; ++f) {
            const bool less = lt(a[f], x);
            rank += less;
            dup = dup || (f != e && !less && !lt(x, a[f]));
        }
        tmp[rank] = x;
    }
    dup = __any_sync(FULLMASK, dup);
    __syncwarp();
    if (dup) { if (lane == 0) introsort((long)n, a, lt); }
    else {
        uint32_t *dst = reinterpret_cast<uint32_t *>(a);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tmp);
        const int words = n * (int)(sizeof(AlnReg) / 4);
        for (int k = lane; k < words; k += 32) dst[k] = src[k];
    }
    __syncwarp();
}

// mem_sort_dedup_patch (bwamem.c:441-493) without a query, i.e. as mem_matesw calls it (bwamem_pair.c:175): no patching
__device__ int sort_dedup_warp(const Opt &opt, int n, AlnReg *a, AlnReg *tmp)
{
    const int lane = threadIdx.x & 31;
    if (n <= 1) return n;
    warp_sort_regs(n, a, tmp, LtRegRe());
    int m = 0;
    if (lane == 0) {
        int i, j;
        for (i = 0; i < n; ++i) a[i].n_comp = 1;
        for (i = 1; i < n; ++i) {
            AlnReg &p = a[i];
            if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt.max_chain_gap) continue;
            for (j = i - 1; j >= 0 && p.rid == a[j].rid && p.rb < a[j].re + opt.max_chain_gap; --j) {
                AlnReg &q = a[j];
                if (q.qe == q.qb) continue;
                const int64_t orr = q.re - p.rb;
                const int64_t oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
                const int64_t mr = q.re - q.rb < p.re - p.rb ? q.re - q.rb : p.re - p.rb;
                const int64_t mq = q.qe - q.qb < p.qe - p.qb ? q.qe - q.qb : p.qe - p.qb;
                if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
                    if (p.score < q.score) { p.qe = p.qb; break; }
                    else q.qe = q.qb;
                }
            }
        }
        for (i = 0, m = 0; i < n; ++i)
            if (a[i].qe > a[i].qb) {
                if (m != i) a[m++] = a[i];
                else ++m;
            }
    }
    n = __shfl_sync(FULLMASK, m, 0);
    __syncwarp();
    warp_sort_regs(n, a, tmp, LtRegScore());
    if (lane == 0) {
        int i;
        for (i = 1; i < n; ++i)
            if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
        for (i = 1, m = 1; i < n; ++i)
            if (a[i].qe > a[i].qb) {
                if (m != i) a[m++] = a[i];
                else ++m;
            }
    }
    m = __shfl_sync(FULLMASK, m, 0);
    __syncwarp();
    return m;
}

// mem_matesw with the Smith-Waterman results at hand; warp-uniform (every lane the same control flow on the same values)
__device__ int mate_rescue_warp_pre(const Opt &opt, const IndexView &ix, const PeStat pes[4], const AlnReg &a, int l_ms, RegList &ma, AlnReg *tmp,
                                    int call_key, const RescuePre &pre, int *err)
{
    const int lane = threadIdx.x & 31;
    const int64_t l_pac = ix.l_pac;
    int i, r, skip[4], n = 0;
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
            int k = 0;
            while (k < pre.n && pre.jobs[k].key != (call_key | r)) ++k;
            if (k == pre.n) { *err = ERR_SCRATCH_OVERFLOW; return n; }
            const SwResult aln = pre.res[k];
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
                __syncwarp();
                if (lane == 0) {
                    for (i = 0; i < ma.n - 1; ++i)
                        if (ma.a[i].score < b.score) break;
                    const int at = i;
                    for (i = ma.n - 1; i > at; --i) ma.a[i] = ma.a[i - 1];
                    ma.a[i] = b;
                }
                __syncwarp();
            }
            ++n;
        }
        if (n) ma.n = sort_dedup_warp(opt, ma.n, ma.a, tmp);
    }
    return n;
}

// One queued pair on one warp: the rescue block of mem_sam_pe (bwamem_pair.c:262-277) replayed over the results, then the rest
// of the pair's finalisation on lane 0. lists: 2 * (ws.reg_cap + max_matesw) regions for the two ends + ws.reg_cap for the sort.
__device__ void stage_final_pe_replay_warp(const Opt &opt, const IndexView &ix, BatchDev &B, int p, FinalWS &ws, AlnReg *lists, const RescuePre &pre)
{
    const int lane = threadIdx.x & 31;
    const int r0 = p << 1, r1 = r0 | 1;
    const int stride = ws.reg_cap + opt.max_matesw;
    AlnReg *tmp = lists + (size_t)2 * stride;
    const int e0 = B.err[r0] ? B.err[r0] : B.err[r1];
    if (lane == 0) { readout_init(B.out[r0], e0); readout_init(B.out[r1], e0); }
    __syncwarp();
    if (e0) return;
    RegList rl[2];
    for (int i = 0; i < 2; ++i) {
        const int r = r0 | i, n = B.n_regs[r];
        rl[i].a = lists + (size_t)i * stride; rl[i].cap = ws.reg_cap; rl[i].n = n;
        if (n > ws.reg_cap) { if (lane == 0) B.out[r0].err = B.out[r1].err = ERR_SCRATCH_OVERFLOW; return; }
        const uint32_t *src = reinterpret_cast<const uint32_t *>(B.regs + B.seed_off[r]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(rl[i].a);
        for (int k = lane; k < n * (int)(sizeof(AlnReg) / 4); k += 32) dst[k] = src[k];
    }
    __syncwarp();
    int err = 0;
    const int ls[2] = {(int)(B.seq_off[r0 + 1] - B.seq_off[r0]), (int)(B.seq_off[r1 + 1] - B.seq_off[r1])};
    const uint8_t *seqs[2] = {B.seq + B.seq_off[r0], B.seq + B.seq_off[r1]};
    AlnReg *bcopy[2]; int nb[2];
    for (int i = 0; i < 2; ++i) {
        nb[i] = 0;
        bcopy[i] = rl[i].a + rl[i].cap;
        for (int j = 0; j < rl[i].n; ++j)
            if (rl[i].a[j].score >= rl[i].a[0].score - opt.pen_unpaired) {
                if (nb[i] < opt.max_matesw && lane == 0) bcopy[i][nb[i]] = rl[i].a[j];
                ++nb[i];
            }
        if (nb[i] > opt.max_matesw) nb[i] = opt.max_matesw;
    }
    __syncwarp();
    for (int i = 0; i < 2 && !err; ++i)
        for (int j = 0; j < nb[i] && !err; ++j) {
            const AlnReg a = bcopy[i][j];
            mate_rescue_warp_pre(opt, ix, B.pes, a, ls[!i], rl[!i], tmp, i << 16 | j << 2, pre, &err);
        }
    __syncwarp();
    if (lane == 0) {
        if (!err) {
            Opt o2 = opt;
            o2.flag |= F_NO_RESCUE;
            finalize_pair(o2, ix, B.mt, B.pes, (uint64_t)((B.n_processed >> 1) + p), r0, ls[0], seqs[0], rl[0], ls[1], seqs[1], rl[1],
                          ws, B.arena, B.tasks, B.out, &err);
        }
        if (err) B.out[r0].err = B.out[r1].err = err;
    }
    __syncwarp();
}

// One pair that needs rescue Smith-Waterman, on one warp: the rescue block of mem_sam_pe (bwamem_pair.c:262-277) with
// the lanes together, then the rest of the pair's finalisation on lane 0 with the rescue already done.
__device__ void stage_final_pe_heavy(const Opt &opt, const IndexView &ix, BatchDev &B, int p, FinalWS &ws, AlnReg *wregs, const WarpSw &sw)
{
    const int lane = threadIdx.x & 31;
    const int r0 = p << 1, r1 = r0 | 1;
    const int stride = ws.reg_cap + opt.max_matesw;
    RegList rl[2];
    for (int i = 0; i < 2; ++i) {
        const int r = r0 | i, n = B.n_regs[r];
        rl[i].a = wregs + (size_t)i * stride; rl[i].cap = ws.reg_cap; rl[i].n = n;
        const AlnReg *src = B.regs + B.seed_off[r];
        for (int j = lane; j < n; j += 32) rl[i].a[j] = src[j];
    }
    __syncwarp();
    int err = 0;
    const int ls[2] = {(int)(B.seq_off[r0 + 1] - B.seq_off[r0]), (int)(B.seq_off[r1 + 1] - B.seq_off[r1])};
    const uint8_t *seqs[2] = {B.seq + B.seq_off[r0], B.seq + B.seq_off[r1]};
    AlnReg *bcopy[2]; int nb[2];
    for (int i = 0; i < 2; ++i) {
        nb[i] = 0;
        bcopy[i] = rl[i].a + rl[i].cap;
        for (int j = 0; j < rl[i].n; ++j)
            if (rl[i].a[j].score >= rl[i].a[0].score - opt.pen_unpaired) {
                if (nb[i] < opt.max_matesw && lane == 0) bcopy[i][nb[i]] = rl[i].a[j];
                ++nb[i];
            }
        if (nb[i] > opt.max_matesw) nb[i] = opt.max_matesw;
    }
    __syncwarp();
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < nb[i]; ++j) {
            const AlnReg a = bcopy[i][j];
            mate_rescue_warp(opt, ix, B.pes, a, ls[!i], seqs[!i], rl[!i], ws, sw, &err);
        }
    __syncwarp();
    if (lane == 0) {
        Opt o2 = opt;
        o2.flag |= F_NO_RESCUE;
        finalize_pair(o2, ix, B.mt, B.pes, (uint64_t)((B.n_processed >> 1) + p), r0, ls[0], seqs[0], rl[0], ls[1], seqs[1], rl[1],
                      ws, B.arena, B.tasks, B.out, &err);
        if (err) B.out[r0].err = B.out[r1].err = err;
    }
    __syncwarp();
}

} // namespace bsb
