// bsb_index.h -- the HBM-resident index view and the FM-index / reference-fetch primitives.
//
// Layout in HBM (identical to the reference's on-disk/in-RAM layout):
//   bwt   : one 64-byte block per 128 BWT symbols = 4 x u64 cumulative A/C/G/T counts followed by
//           8 x u32 packed symbols (symbol i of a word at bits 30-2*(i&15))   (bwt.h:72-78)
//   sa    : SA sampled every `sa_intv` ranks, u64 entries, sa[0] = -1               (bwt.c:61-84)
//   pac   : converted reference, 2 bit/base, forward strands only: watson(C->T) contigs then
//           crick(G->A) contigs; opac: the same coordinates, unconverted bases    (bntseq.c:239-361)
// A block is 64-byte aligned, so one occ query is exactly two 32-byte sectors.
#pragma once
#include "bsb_hd.h"

namespace bsb {

struct IndexView {
    const uint32_t *bwt;
    const uint64_t *sa;
    const uint8_t *pac, *opac;
    const Ann *anns;
    uint64_t primary, L2[5], seq_len;
    int64_t l_pac, crick_l;
    int32_t n_seqs, sa_intv;
    // optional result-preserving denser SA (values are independent of the sampling rate)
    const uint32_t *sa32;   // when non-null: SA sampled every sa32_intv ranks, low 32 bits of each entry
    int32_t sa32_intv, pad_;
    const uint8_t *sa_hi;   // bits 32..39 of the sa32 entries (40-bit SA for texts of >= 2^32 symbols: 5 bytes per rank,
                            // 62 GB for a human-scale 12.4 G-symbol text -- resident in 180 GB of HBM); null when < 2^32
    // optional sector-sized occ blocks (seq_len < 2^32): one 32-byte block per 64 BWT symbols = 4 x u32 cumulative
    // counts + 4 packed words, so that one rank query is ONE 32-byte sector and one 256-bit load (occ32_* below)
    const uint32_t *occ32;
};

// ---- occurrence counting -------------------------------------------------------------------

// per-symbol match mask of one packed word: bit 2j set iff symbol j (from the LSB pair) == c
BSB_HD uint32_t sym_eq_mask(uint32_t w, int c)
{
    uint32_t x = w ^ (0x55555555u * (uint32_t)c); // equal symbols become 00
    return ~(x | (x >> 1)) & 0x55555555u;
}

BSB_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

struct OccBlock { uint64_t cnt[4]; uint32_t w[8]; };

BSB_HD void load_block(const uint32_t *bwt, uint64_t blk, OccBlock &b)
{
#if defined(__CUDA_ARCH__)
    // two 256-bit loads (LDG.E.256, sm_100): header and packed symbols, each one 32-byte sector
    const uint32_t *p = bwt + (blk << 4);
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(b.cnt[0]), "=l"(b.cnt[1]), "=l"(b.cnt[2]), "=l"(b.cnt[3]) : "l"(p));
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7]) : "l"(p + 8));
#else
    const uint32_t *p = bwt + (blk << 4);
    for (int i = 0; i < 4; ++i) b.cnt[i] = (uint64_t)p[2 * i + 1] << 32 | p[2 * i];
    for (int i = 0; i < 8; ++i) b.w[i] = p[8 + i];
#endif
}

BSB_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// counts of A,C,G,T in BWT[block_start .. block_start + r] (r in [0,127]), added to the block's
// cumulative counts. Restates bwt_occ4 (bwt.c:169-186) with popcounts instead of the byte table:
// two packed words at a time, C/G/T by popcount, A as the remainder.
BSB_HD void block_occ4(const OccBlock &b, int r, uint64_t cnt[4])
{
    const int nsym = r + 1;                 // symbols to count
    uint32_t c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // symbols 32*i .. 32*i+31: word 2i holds the first 16 (high half of the 64-bit view)
        uint64_t w = (uint64_t)b.w[2 * i] << 32 | b.w[2 * i + 1];
        int k = nsym - 32 * i;              // how many of these 32 symbols are wanted
        k = k < 0 ? 0 : k > 32 ? 32 : k;
        uint64_t keep = k == 0 ? 0ull : (~0ull << ((32 - k) << 1));
        keep &= 0x5555555555555555ull;
        uint64_t hi = (w >> 1), lo = w;
        c1 += popc64(~hi & lo & keep);
        c2 += popc64(hi & ~lo & keep);
        c3 += popc64(hi & lo & keep);
    }
    cnt[0] = b.cnt[0] + (uint32_t)nsym - c1 - c2 - c3; cnt[1] = b.cnt[1] + c1; cnt[2] = b.cnt[2] + c2; cnt[3] = b.cnt[3] + c3;
}

BSB_HD void occ4(const IndexView &ix, uint64_t k, uint64_t cnt[4])
{
    if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    k -= (k >= ix.primary);
    OccBlock b;
    load_block(ix.bwt, k >> 7, b);
    block_occ4(b, (int)(k & 127), cnt);
}

// bwt_2occ4 (bwt.c:189-220). Branch-free on the device: both blocks are always fetched (the second fetch
// hits the same line when k and l share a block), so the lanes of a warp never split here.
BSB_HD void occ4_pair(const IndexView &ix, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4])
{
    const bool kz = k == (uint64_t)-1, lz = l == (uint64_t)-1;
    const uint64_t _k = kz ? 0 : k - (k >= ix.primary), _l = lz ? 0 : l - (l >= ix.primary);
    OccBlock bk, bl;
    load_block(ix.bwt, _k >> 7, bk);
    load_block(ix.bwt, _l >> 7, bl);
    block_occ4(bk, (int)(_k & 127), ck);
    block_occ4(bl, (int)(_l & 127), cl);
    if (kz) ck[0] = ck[1] = ck[2] = ck[3] = 0;
    if (lz) cl[0] = cl[1] = cl[2] = cl[3] = 0;
}

// bwt_extend (bwt.c:262-275)
BSB_HD void fm_extend(const IndexView &ix, const Intv &ik, Intv ok[4], int is_back)
{
    uint64_t tk[4], tl[4];
    uint64_t xa = is_back ? ik.x0 : ik.x1; // x[!is_back]
    uint64_t xb = is_back ? ik.x1 : ik.x0; // x[is_back]
    occ4_pair(ix, xa - 1, xa - 1 + ik.x2, tk, tl);
    uint64_t na[4], sz[4], nb[4];
    for (int i = 0; i < 4; ++i) {
        na[i] = ix.L2[i] + 1 + tk[i];
        sz[i] = tl[i] - tk[i];
    }
    nb[3] = xb + (xa <= ix.primary && xa + ik.x2 - 1 >= ix.primary);
    nb[2] = nb[3] + sz[3];
    nb[1] = nb[2] + sz[2];
    nb[0] = nb[1] + sz[1];
    for (int i = 0; i < 4; ++i) {
        if (is_back) { ok[i].x0 = na[i]; ok[i].x1 = nb[i]; }
        else { ok[i].x1 = na[i]; ok[i].x0 = nb[i]; }
        ok[i].x2 = sz[i];
    }
}

BSB_HD uint64_t sel4(int c, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3)
{
    return c == 0 ? a0 : c == 1 ? a1 : c == 2 ? a2 : a3;
}

// fm_extend() when only the interval of symbol c is wanted; no indexable temporaries, so everything stays
// in registers on the device
BSB_HD Intv fm_extend_sel(const IndexView &ix, const Intv &ik, int c, int is_back)
{
    uint64_t tk[4], tl[4];
    const uint64_t xa = is_back ? ik.x0 : ik.x1, xb = is_back ? ik.x1 : ik.x0;
    occ4_pair(ix, xa - 1, xa - 1 + ik.x2, tk, tl);
    const uint64_t s0 = tl[0] - tk[0], s1 = tl[1] - tk[1], s2 = tl[2] - tk[2], s3 = tl[3] - tk[3];
    const uint64_t n3 = xb + (xa <= ix.primary && xa + ik.x2 - 1 >= ix.primary);
    const uint64_t n2 = n3 + s3, n1 = n2 + s2, n0 = n1 + s1;
    const uint64_t na = sel4(c, ix.L2[0] + 1 + tk[0], ix.L2[1] + 1 + tk[1], ix.L2[2] + 1 + tk[2], ix.L2[3] + 1 + tk[3]);
    const uint64_t nb = sel4(c, n0, n1, n2, n3);
    Intv o;
    if (is_back) { o.x0 = na; o.x1 = nb; } else { o.x1 = na; o.x0 = nb; }
    o.x2 = sel4(c, s0, s1, s2, s3);
    o.info = 0;
    return o;
}

BSB_HD void fm_set_intv(const IndexView &ix, int c, Intv &ik)
{
    ik.x0 = ix.L2[c] + 1;
    ik.x2 = ix.L2[c + 1] - ix.L2[c];
    ik.x1 = ix.L2[3 - c] + 1;
    ik.info = 0;
}

// ---- sector-sized occ blocks -----------------------------------------------------------------
// Same BWT, same counts as the reference blocks (bwt.h:72-78), re-cut at index load: block b32 covers symbols
// [64*b32, 64*b32 + 64); its counts are the reference block's, plus the first 64 symbols of that block for odd b32.
BSB_HD void occ32_make_block(const uint32_t *bwt, uint64_t b32, uint32_t out[8])
{
    const uint32_t *p = bwt + ((b32 >> 1) << 4);
    uint64_t cnt[4];
    for (int j = 0; j < 4; ++j) cnt[j] = (uint64_t)p[2 * j + 1] << 32 | p[2 * j];
    const int half = (int)(b32 & 1);
    if (half)
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 4; ++c) cnt[c] += (uint64_t)popc32(sym_eq_mask(p[8 + i], c));
    for (int j = 0; j < 4; ++j) out[j] = (uint32_t)cnt[j];
    for (int i = 0; i < 4; ++i) out[4 + i] = p[8 + 4 * half + i];
}

struct Occ32 { uint32_t cnt[4], w[4]; };

BSB_HD void occ32_load(const uint32_t *occ32, uint64_t blk, Occ32 &b)
{
    const uint32_t *p = occ32 + (blk << 3);
#if defined(__CUDA_ARCH__)
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(b.cnt[0]), "=r"(b.cnt[1]), "=r"(b.cnt[2]), "=r"(b.cnt[3]), "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]) : "l"(p));
#else
    for (int i = 0; i < 4; ++i) { b.cnt[i] = p[i]; b.w[i] = p[4 + i]; }
#endif
}

// mask of the first k (0..16) symbols of a packed word, at the low bit of each symbol
BSB_HD uint32_t occ32_keep(int k)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(0u, 0x55555555u, (unsigned)(2 * k));
#else
    return k == 0 ? 0u : 0x55555555u << (32 - 2 * k);
#endif
}

// E = occ(c, p), G = sum of occ(j, p) over j > c, for the position p = 64*blk + r of block b. With P1/P2/P3 = number
// of symbols >= 1 / >= 2 / >= 3 among the first r+1 of the block: symbols >= t form the sequence (r+1, P1, P2, P3, 0).
BSB_HD void occ32_eg(const Occ32 &b, int r, int c, uint32_t &E, uint32_t &G)
{
    const int nsym = r + 1;
    uint32_t a[4], h[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = nsym - 16 * i;
        k = k < 0 ? 0 : k > 16 ? 16 : k;
        const uint32_t keep = occ32_keep(k), hi = b.w[i] >> 1, lo = b.w[i];
        a[i] = (hi | lo) & keep; h[i] = hi & keep; d[i] = hi & lo & keep;
    }
    const uint32_t P1 = popc32(a[0] | a[1] << 1) + popc32(a[2] | a[3] << 1);
    const uint32_t P2 = popc32(h[0] | h[1] << 1) + popc32(h[2] | h[3] << 1);
    const uint32_t P3 = popc32(d[0] | d[1] << 1) + popc32(d[2] | d[3] << 1);
    uint32_t ge_c = (uint32_t)nsym, ge_c1 = P1, he = b.cnt[0], hg = b.cnt[1] + b.cnt[2] + b.cnt[3];
    if (c >= 1) { ge_c = P1; ge_c1 = P2; he = b.cnt[1]; hg = b.cnt[2] + b.cnt[3]; }
    if (c >= 2) { ge_c = P2; ge_c1 = P3; he = b.cnt[2]; hg = b.cnt[3]; }
    if (c >= 3) { ge_c = P3; ge_c1 = 0; he = b.cnt[3]; hg = 0; }
    E = he + ge_c - ge_c1; G = hg + ge_c1;
}

// fm_extend_one (bsb_seed3.h) over the sector-sized blocks; every coordinate is < 2^32 (seq_len + 1 < 2^32)
BSB_HD void fm_extend_one32(const IndexView &ix, uint32_t xa, uint32_t xb, uint32_t s, int c, uint32_t &na, uint32_t &nb, uint32_t &sz)
{
    const uint32_t primary = (uint32_t)ix.primary;
    const uint32_t k = xa - 1, l = xa - 1 + s;
    const bool kz = k == 0xffffffffu, lz = l == 0xffffffffu;
    const uint32_t _k = kz ? 0 : k - (k >= primary), _l = lz ? 0 : l - (l >= primary);
    Occ32 bk, bl;
    occ32_load(ix.occ32, _k >> 6, bk);
    occ32_load(ix.occ32, _l >> 6, bl);
    uint32_t ek, gk, el, gl;
    occ32_eg(bk, (int)(_k & 63), c, ek, gk);
    occ32_eg(bl, (int)(_l & 63), c, el, gl);
    if (kz) ek = gk = 0;
    if (lz) el = gl = 0;
    na = (uint32_t)ix.L2[c] + 1 + ek;
    sz = el - ek;
    nb = xb + (xa <= primary && xa + s - 1 >= primary) + (gl - gk);
}

// ---- suffix-array lookup -------------------------------------------------------------------

// one LF step: bwt_invPsi (bwt.c:53-59) with the symbol fetch and the rank sharing one block load
BSB_HD uint64_t fm_lf(const IndexView &ix, uint64_t k)
{
    if (k == ix.primary) return 0;
    uint64_t x = k - (k > ix.primary);
    OccBlock b;
    load_block(ix.bwt, x >> 7, b);
    int r = (int)(x & 127);
    int c = (b.w[r >> 4] >> ((~r & 15) << 1)) & 3;
    // rank of c in BWT[0..x]
    uint32_t n = 0;
    int nfull = r >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i <= nfull) {
            uint32_t m = sym_eq_mask(b.w[i], c);
            if (i == nfull) m &= ~((1u << ((~r & 15) << 1)) - 1u);
            n += popc32(m);
        }
    }
    return ix.L2[c] + b.cnt[c] + n;
}

// bwt_sa (bwt.c:86-96)
BSB_HD uint64_t fm_sa(const IndexView &ix, uint64_t k)
{
    uint64_t steps = 0;
    if (ix.sa32) {
        uint64_t mask = (uint64_t)ix.sa32_intv - 1;
        while (k & mask) { ++steps; k = fm_lf(ix, k); }
        uint64_t j = k / (uint64_t)ix.sa32_intv;
        uint64_t v = j == 0 ? (uint64_t)-1 : (uint64_t)ix.sa32[j];
        if (ix.sa_hi && j != 0) v |= (uint64_t)ix.sa_hi[j] << 32;
        return steps + v;
    }
    uint64_t mask = (uint64_t)ix.sa_intv - 1;
    while (k & mask) { ++steps; k = fm_lf(ix, k); }
    return steps + ix.sa[k / (uint64_t)ix.sa_intv];
}

// ---- reference coordinates & sequence --------------------------------------------------------

BSB_HD int pac_get(const uint8_t *pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

// base at doubled coordinate p in [0, 2*l_pac): reverse half is the reverse complement (bntseq.c:413-434)
BSB_HD int ref_base(const uint8_t *pac, int64_t l_pac, int64_t p)
{
    return p < l_pac ? pac_get(pac, p) : 3 - pac_get(pac, (l_pac << 1) - 1 - p);
}

BSB_HD int64_t depos(int64_t l_pac, int64_t pos, int *is_rev)
{
    return (*is_rev = (pos >= l_pac)) ? (l_pac << 1) - 1 - pos : pos;
}

// bns_pos2rid (bntseq.c:364-378)
BSB_HD int pos2rid(const IndexView &ix, int64_t pos_f)
{
    if (pos_f >= ix.l_pac) return -1;
    int left = 0, mid = 0, right = ix.n_seqs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= ix.anns[mid].offset) {
            if (mid == ix.n_seqs - 1) break;
            if (pos_f < ix.anns[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

// bns_intv2rid (bntseq.c:380-388)
BSB_HD int intv2rid(const IndexView &ix, int64_t rb, int64_t re)
{
    int is_rev;
    if (rb < ix.l_pac && re > ix.l_pac) return -2;
    int rid_b = pos2rid(ix, depos(ix.l_pac, rb, &is_rev));
    int rid_e = rb < re ? pos2rid(ix, depos(ix.l_pac, re - 1, &is_rev)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

// bns_fetch_seq (bntseq.c:436-461) without materialising the sequence: clamps [beg,end) to the
// contig (and strand) that contains mid and returns its id. Bases are then read with ref_base().
BSB_HD int fetch_window(const IndexView &ix, int64_t *beg, int64_t mid, int64_t *end)
{
    if (*end < *beg) tswap(*beg, *end);
    int is_rev;
    int rid = pos2rid(ix, depos(ix.l_pac, mid, &is_rev));
    int64_t far_beg = ix.anns[rid].offset;
    int64_t far_end = far_beg + ix.anns[rid].len;
    if (is_rev) {
        int64_t tmp = far_beg;
        far_beg = (ix.l_pac << 1) - far_end;
        far_end = (ix.l_pac << 1) - tmp;
    }
    *beg = *beg > far_beg ? *beg : far_beg;
    *end = *end < far_end ? *end : far_end;
    return rid;
}

} // namespace bsb
