// bsb_deflate.h -- one BGZF block (RFC 1951 deflate, dynamic Huffman codes, inside the gzip member htslib's bgzf.c writes)
// produced by ONE thread block from up to 0xff00 bytes of BAM records.
//
// Replaces the compression half of `stream_bam` (bsbolt/External/HTSLIB/stream_bam.c -> bgzf_write -> deflate, zlib on the
// host cores): with the records already in HBM (k_bam_write) the compressed blocks are what crosses PCIe and the host only
// copies them to the file. Only the *uncompressed* stream is comparable with the reference (tests: inflate and compare);
// the compressed bytes are this file's own.
//
// Shape (everything data-parallel over the block's threads, nothing is a per-thread zlib; the input block is staged in
// shared memory first, and nothing a thread indexes dynamically lives in local memory -- with ~200 KB of the SM's L1 carved
// out as shared memory a local access is an L2 round trip):
//   1. Matches. The block is taken in chunks of DF_CH = 512 positions, one per thread. Every position hashes its next 4
//      bytes and looks up the most recent earlier-chunk position with that hash (a 16-bit atomic-max table in shared memory,
//      so the table and hence the output do not depend on thread timing; of neighbours with the same hash -- a run -- only the
//      last one inserts) and compares 8 bytes a step, at most DF_MAX_HASH_MATCH bytes. The distance-1 candidate (runs) costs
//      no compare: one bit per position says in[p] == in[p-1] and a run's length is a count of one bits. The longer wins.
//   2. Parse. Greedy: from the position where the previous token ended, take the match there (or one literal) and jump
//      behind it. The positions visited are found by pointer doubling over the chunk's jump table instead of a serial walk:
//      five rounds INSIDE each warp's 32 positions (warp barriers only) give every position's way out of its segment,
//      16 look-ups chain the segments, five more warp-local rounds mark the next 2^k hops; marked positions are compacted
//      into tokens (prefix over the mark words) and counted into the literal/length and distance histograms.
//   3. Codes. Symbols are rank-sorted by frequency in parallel; one thread per tree runs the two-queue Huffman merge (the
//      only serial pass left); depths, the Kraft repair of the depth limit, the assignment of lengths and the canonical codes
//      run across the block. The header: one thread per RUN of equal code lengths emits its run-length symbols (count, scan,
//      write), one thread builds the 19-symbol code, then every symbol is sized, placed by a scan and ORed into the header.
//   4. Bits. Every thread sizes a contiguous run of tokens, a scan gives its first bit, it packs its run in a 64-bit
//      register and ORs whole words into the (zeroed) output; CRC-32 of the input is computed per slice and combined
//      with x^n mod P multiplications.
// A block that does not shrink is stored (BTYPE 00).
//
// The functions take an execution policy X: X::par(n, f) runs f(i) for i in [0, n) across the block and ends with a barrier
// (X::par1(f): n = DF_CH, one item per thread; X::wpar1(f): the same with a WARP barrier, for phases in which a thread only
// reads what its own warp's 32 threads wrote since the last block barrier; X::sync(): a block barrier; X::tick(k): a
// measurement hook, cycles since the previous tick are charged to phase k);
// X::atomic_* are the shared/global atomics. XDev (bsb_cuda.cu) maps them to threadIdx/__syncthreads/atomicOr; XHost
// (tests/hostsim/bamsim.cpp) runs the same phases as plain loops, which is how this file is tested on the CPU against zlib's
// inflate and crc32 -- a phase therefore never reads what the same phase writes, except through the atomics.
#pragma once
#include <stdint.h>
#include <string.h>
#include "bsb_hd.h"

namespace bsb {

constexpr int BGZF_MAX_IN = 0xff00;          // htslib BGZF_BLOCK_SIZE
constexpr int BGZF_SLOT = 0x10000;           // room for one block in the worst case: 18 + 5 + 0xff00 + 8
constexpr int DF_CH = 512;                   // positions per chunk = threads per block
constexpr int DF_HASH_BITS = 13;
constexpr int DF_MIN_MATCH = 4, DF_MAX_MATCH = 258, DF_MAX_DIST = 32768, DF_GOOD_RUN = 32;
constexpr int DF_MAX_HASH_MATCH = 128;         // a match found through the hash table is compared this far (runs go to 258 through the eq bits): bounds the slowest thread
constexpr int DF_NLL = 286, DF_ND = 30, DF_NCL = 19;

struct DeflateShared {
    uint32_t in32[BGZF_MAX_IN / 4 + 4];      // the block's input: every later access (hashing, match compares, literals, CRC) is a shared-memory access
    uint32_t head[(1 << DF_HASH_BITS) / 2];  // hash -> 1 + the latest position of an earlier chunk (0: none), two 16-bit entries per word
    uint32_t eq[BGZF_MAX_IN / 32 + 12];       // bit p: in[p] == in[p - 1] -- a run's length is a count of one bits, not a byte compare
    uint16_t mlen[DF_CH], mdist[DF_CH], hash[DF_CH];
    uint16_t jl[6][DF_CH];                    // jl[k][t]: where 2^k hops from t lead, hops that stay inside t's 32-position segment
    uint32_t markw[DF_CH / 32];
    uint32_t hist[320];                      // [0, 286): literal/length, [288, 318): distance
    uint32_t sorted[320];                    // (freq << 9 | symbol) ascending, per tree at the same bases
    uint32_t wi[320];                        // two-queue merge: weights of the internal nodes
    uint16_t par_leaf[320], par_int[320], depth_int[320];
    uint8_t len[320];                        // code lengths
    uint16_t code[320];                      // codes, bit-reversed (deflate sends Huffman codes MSB first in an LSB-first stream)
    uint16_t rle[320];                       // run-length coded code lengths: symbol | extra << 8
    uint32_t hdr[80];                        // the block header's bits
    uint32_t part[DF_CH + 1];
    // what one thread's serial passes index dynamically lives here, not in local memory (with ~200 KB of the SM's L1 carved out
    // as shared memory, a local-memory access is an L2 round trip)
    int32_t entry[DF_CH / 32 + 1]; uint32_t pre[DF_CH / 32 + 1];
    uint32_t tscr[3][52];                     // df_tree: code counts per length, next code per length (literal/length, distance, code-length tree)
    uint32_t clf[DF_NCL], cl_sorted[DF_NCL]; uint8_t cl_len[DF_NCL + 1], order[DF_NCL + 1]; uint16_t cl_code[DF_NCL + 1];
    uint32_t crc_tab[256], x2n[32];
    uint32_t crc;
    int n_used[2], hdr_bits, stored, hlit, hdist;
    uint16_t rcnt[320], roff[320]; uint32_t rgrp[11];   // the header's run-length symbols: counts / bit sizes, their prefix sums
    uint32_t total_bits;
};

// ---------------------------------------------------------------------------------------------------------------------
BSB_HD uint32_t df_load32(const uint8_t *p)                       // four bytes at any address, little-endian
{
#if defined(__CUDA_ARCH__)
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) << 3;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
#else
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
#endif
}
// the same out of an array of little-endian words (the staged input): byte offset `at`
BSB_HD uint32_t df_word(const uint32_t *w, int at)
{
    const uint32_t lo = w[at >> 2], hi = w[(at >> 2) + 1];
    const uint32_t sh = (uint32_t)(at & 3) << 3;
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? lo >> sh | hi << (32 - sh) : lo;
#endif
}
BSB_HD uint32_t df_byte(const uint32_t *w, int at) { return w[at >> 2] >> ((at & 3) << 3) & 0xffu; }
BSB_HD int df_ctz(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
BSB_HD int df_hibit(uint32_t x)                                   // position of the highest set bit, x > 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
BSB_HD int df_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

BSB_HD uint32_t df_funnel(uint32_t lo, uint32_t hi, uint32_t sh)     // bits [sh, sh + 32) of hi:lo, sh in {0, 8, 16, 24}
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? lo >> sh | hi << (32 - sh) : lo;
#endif
}
// bytes at offsets a.. and b.. of the staged input agree for how long (at most lim). Eight bytes a step, all six loads of a
// step issued together: the loop is a chain of shared-memory latencies and the slowest thread of the block sets the pace.
BSB_HD int df_match_len(const uint32_t *w, int a, int b, int lim)
{
    int k = 0;
    const uint32_t sa = (uint32_t)(a & 3) << 3, sb = (uint32_t)(b & 3) << 3;
    for (; k + 8 <= lim; k += 8) {
        const int ia = (a + k) >> 2, ib = (b + k) >> 2;
        const uint32_t a0 = w[ia], a1 = w[ia + 1], a2 = w[ia + 2], b0 = w[ib], b1 = w[ib + 1], b2 = w[ib + 2];
        const uint32_t x0 = df_funnel(a0, a1, sa) ^ df_funnel(b0, b1, sb), x1 = df_funnel(a1, a2, sa) ^ df_funnel(b1, b2, sb);
        if (x0) return k + (df_ctz(x0) >> 3);
        if (x1) return k + 4 + (df_ctz(x1) >> 3);
    }
    for (; k + 4 <= lim; k += 4) {
        const uint32_t x = df_word(w, a + k) ^ df_word(w, b + k);
        if (x) return k + (df_ctz(x) >> 3);
    }
    while (k < lim && df_byte(w, a + k) == df_byte(w, b + k)) ++k;
    return k;
}
// the distance-1 match at i: how many one bits follow from bit i of eq[] (at most lim; eq[] is zero behind the input)
BSB_HD int df_run_len(const uint32_t *eq, int i, int lim)
{
    int w = i >> 5, got = 0, avail = 32 - (i & 31);
    uint32_t z = ~eq[w] >> (i & 31);            // bit k set: eq bit i + k is zero; only the low `avail` bits are this word's
    for (;;) {
        if (z) { const int f = df_ctz(z); return got + f < lim ? got + f : lim; }
        got += avail;
        if (got >= lim) return lim;
        z = ~eq[++w]; avail = 32;
    }
}

// length 3..258 -> literal/length symbol, number and value of the extra bits (RFC 1951 3.2.5)
BSB_HD void df_len_code(int l, int &sym, int &nx, int &xv)
{
    if (l == 258) { sym = 285; nx = 0; xv = 0; return; }
    const int m = l - 3;
    if (m < 8) { sym = 257 + m; nx = 0; xv = 0; return; }
    const int hb = df_hibit((uint32_t)m);
    sym = 257 + 4 * (hb - 1) + ((m >> (hb - 2)) & 3);
    nx = hb - 2; xv = m & ((1 << nx) - 1);
}
// distance 1..32768 -> distance symbol, extra bits
BSB_HD void df_dist_code(int d, int &sym, int &nx, int &xv)
{
    if (d <= 4) { sym = d - 1; nx = 0; xv = 0; return; }
    const int m = d - 1, hb = df_hibit((uint32_t)m);
    sym = 2 * hb + ((m >> (hb - 1)) & 1);
    nx = hb - 1; xv = m & ((1 << nx) - 1);
}

BSB_HD uint32_t df_rev(uint32_t c, int n)
{
#if defined(__CUDA_ARCH__)
    return n ? __brev(c) >> (32 - n) : 0u;
#else
    uint32_t r = 0; for (int k = 0; k < n; ++k) { r = r << 1 | (c & 1); c >>= 1; } return r;
#endif
}

// CRC-32 arithmetic over GF(2), reflected polynomial 0xedb88320: a(x) * b(x) mod P(x)
BSB_HD uint32_t df_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
// x^(8 n) mod P from the table of x^(2^k) mod P
BSB_HD uint32_t df_x8n(const uint32_t *x2n, uint32_t n)
{
    uint32_t p = 0x80000000u;                 // x^0
    for (int k = 3; n; n >>= 1, ++k)
        if (n & 1) p = df_mulmod(x2n[k & 31], p);
    return p;
}

// ---------------------------------------------------------------------------------------------------------------------
// Code lengths of one tree, in pieces so that the block can run the data-parallel ones across its threads.
// sorted[0, m): (freq << 9 | symbol) ascending. scr: 52 words -- [0, 16) codes per length, [16, 33) next canonical code per
// length, [33, 50) rank in `sorted` at which each length starts.
//
// (1) serial, m >= 2: the Huffman merge with two queues -- the leaves in ascending weight, the internal nodes in the order
// they are made (ascending too); on a tie the leaf goes first, which keeps the tree shallow -- and the depth of every internal node
BSB_HD void df_tree_merge(const uint32_t *sorted, int m, uint32_t *wi, uint16_t *par_leaf, uint16_t *par_int, uint16_t *depth_int)
{
    // the weights at the heads of the two queues wait in registers (the loop is one thread's chain of shared-memory latencies);
    // NONE: that queue is empty -- for the internal nodes also "the next one is being made in this step"
    const uint32_t NONE = 0xffffffffu;
    int a = 0, b = 0;
    uint32_t lw = sorted[0] >> 9, iw = NONE;
    for (int j = 0; j < m - 1; ++j) {
        uint32_t w = 0;
        for (int k = 0; k < 2; ++k) {
            if (lw != NONE && lw <= iw) { w += lw; par_leaf[a++] = (uint16_t)j; lw = a < m ? sorted[a] >> 9 : NONE; }
            else { w += iw; par_int[b++] = (uint16_t)j; iw = b < j ? wi[b] : NONE; }
        }
        wi[j] = w;
        if (b == j) iw = w;
    }
    depth_int[m - 2] = 0;
    for (int j = m - 3; j >= 0; --j) depth_int[j] = (uint16_t)(depth_int[par_int[j]] + 1);
}
// (2) per leaf: its code length, clamped to the limit
BSB_HD int df_tree_leaf_len(int m, int i, int limit, const uint16_t *par_leaf, const uint16_t *depth_int)
{
    if (m == 1) return 1;
    const int d = depth_int[par_leaf[i]] + 1;
    return d < limit ? d : limit;
}
// (3) serial, short: the clamped lengths over-subscribe the code space; every step gives one unit of it back (one code leaves
// the last level, one code one level up takes a sibling with it). Then where each length starts among the sorted symbols (the
// rarest take the longest codes) and the first canonical code of each length (RFC 1951 3.2.2)
BSB_HD void df_tree_limit(uint32_t *scr, int limit)
{
    uint32_t *bl = scr, *next = scr + 16, *start = scr + 33;
    uint32_t total = 0;
    for (int d = 1; d <= limit; ++d) total += bl[d] << (limit - d);
    while (total > (1u << limit)) {
        --bl[limit];
        for (int d = limit - 1; d > 0; --d)
            if (bl[d]) { --bl[d]; bl[d + 1] += 2; break; }
        --total;
    }
    uint32_t at = 0;
    for (int d = 15; d >= 1; --d) { start[d] = at; at += d <= limit ? bl[d] : 0u; }
    uint32_t c = 0;
    next[0] = 0;
    for (int d = 1; d <= 15; ++d) { c = (c + bl[d - 1]) << 1; next[d] = c; }
}
// (4) the length of the i-th sorted symbol
BSB_HD int df_tree_len_of_rank(const uint32_t *scr, int limit, uint32_t i)
{
    const uint32_t *bl = scr, *start = scr + 33;
    for (int d = limit; d >= 1; --d)
        if (i < start[d] + bl[d]) return d;
    return 0;
}
// (5) the canonical code of symbol s: the first code of its length + the number of smaller symbols with that length, bit-reversed
BSB_HD uint32_t df_tree_code(const uint32_t *scr, const uint8_t *len, int s)
{
    const int l = len[s];
    if (!l) return 0;
    uint32_t r = 0;
    for (int u = 0; u < s; ++u) r += len[u] == l;
    return df_rev(scr[16 + l] + r, l);
}
// all of it by one thread (the small tree of the code lengths)
BSB_HD void df_tree(const uint32_t *sorted, int m, int limit, uint8_t *len, uint16_t *code, int n_sym,
                    uint32_t *wi, uint16_t *par_leaf, uint16_t *par_int, uint16_t *depth_int, uint32_t *scr)
{
    for (int s = 0; s < n_sym; ++s) { len[s] = 0; code[s] = 0; }
    for (int d = 0; d < 16; ++d) scr[d] = 0;
    if (m >= 2) df_tree_merge(sorted, m, wi, par_leaf, par_int, depth_int);
    for (int i = 0; i < m; ++i) ++scr[df_tree_leaf_len(m, i, limit, par_leaf, depth_int)];
    df_tree_limit(scr, limit);
    for (int i = 0; i < m; ++i) len[sorted[i] & 511] = (uint8_t)df_tree_len_of_rank(scr, limit, (uint32_t)i);
    for (int s = 0; s < n_sym; ++s) code[s] = (uint16_t)df_tree_code(scr, len, s);
}

struct DfBits {                                // the header's bit writer (one thread): bits collect in a register, whole words leave
    uint32_t *w; int n; uint64_t acc; int fill;
    BSB_HD void put(uint32_t v, int nb)
    {
        acc |= (uint64_t)v << fill; fill += nb; n += nb;
        if (fill >= 32) { *w++ = (uint32_t)acc; acc >>= 32; fill -= 32; }
    }
    BSB_HD void finish() { if (fill) *w++ = (uint32_t)acc; }
};

// token: literal byte, or 1 << 31 | (length - 3) << 16 | (distance - 1)
BSB_HD int df_token_bits(const DeflateShared &S, uint32_t tk)
{
    if (!(tk >> 31)) return S.len[tk];
    int sym, nx, xv, ds, dnx, dxv;
    df_len_code((int)((tk >> 16) & 0xff) + 3, sym, nx, xv);
    df_dist_code((int)(tk & 0xffff) + 1, ds, dnx, dxv);
    return S.len[sym] + nx + S.len[288 + ds] + dnx;
}

// ---------------------------------------------------------------------------------------------------------------------
// One BGZF block. in[0, n), 1 <= n <= BGZF_MAX_IN, at any address and readable up to in[n + 7]; out: a 4-byte aligned slot of
// BGZF_SLOT bytes; tok: BGZF_MAX_IN + 8 words of scratch private to the block. Returns (through S.part[0] after the last
// barrier, and as the function's value in every thread) the number of bytes of the finished block.
template <class X>
BSB_HD uint32_t bgzf_block(X &x, DeflateShared &S, const uint8_t *in, int n, uint8_t *out, uint32_t *tok)
{
    uint32_t *out32 = reinterpret_cast<uint32_t *>(out + 16);   // the deflate stream starts at bit 16 of this word array
    const uint32_t *w32 = S.in32;
    // ---- 0. input and tables ----
    x.par(BGZF_MAX_IN / 4 + 4, [&](int w) { S.in32[w] = 4 * w < n ? df_load32(in + 4 * w) : 0u; });
    x.par(BGZF_MAX_IN / 32 + 12, [&](int w) {                    // eq bits: 32 positions per thread and step
        uint32_t bits = 0;
        const int p0 = 32 * w;
        if (p0 < n) {
            for (int g = 0; g < 8; ++g) {
                const int q = p0 + 4 * g;
                const uint32_t a = w32[q >> 2];
                const uint32_t prev = q ? df_byte(w32, q - 1) : (~a & 0xffu);   // (nothing in front of the first byte)
                const uint32_t d = a ^ (a << 8 | prev);              // byte k is zero where in[q + k] == in[q + k - 1]
                bits |= ((d & 0xffu ? 0u : 1u) | (d & 0xff00u ? 0u : 2u) | (d & 0xff0000u ? 0u : 4u) | (d & 0xff000000u ? 0u : 8u)) << (4 * g);
            }
            if (n - p0 < 32) bits &= (1u << (n - p0)) - 1;
        }
        S.eq[w] = bits;
    });
    x.par((1 << DF_HASH_BITS) / 2, [&](int i) { S.head[i] = 0; });
    x.par(320, [&](int i) {
        S.hist[i] = 0;
        if (i < 256) { uint32_t c = (uint32_t)i; for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0xedb88320u : c >> 1; S.crc_tab[i] = c; }
        if (i == 256) {
            uint32_t p = 0x40000000u;          // x^1
            S.x2n[0] = p;
            for (int k = 1; k < 32; ++k) S.x2n[k] = p = df_mulmod(p, p);
            S.crc = 0; S.stored = 0;
        }
        if (i == 257) {
            const uint8_t order[DF_NCL] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            for (int k = 0; k < DF_NCL; ++k) S.order[k] = order[k];   // (constant indices: no local copy)
        }
    });
    x.tick(0);
    // ---- 1 + 2. matches and the greedy parse, chunk by chunk ----
    // Per chunk: (A) every position's best match and its jump; (B) five doubling rounds INSIDE each 32-position segment --
    // a segment belongs to one warp, so these rounds need a warp barrier only -- leave in jl[5][t] the position at which
    // the chain starting at t leaves t's segment; (C) the entry of each segment follows from the chunk's entry by eight
    // table look-ups; (D) five more warp-local rounds mark the chain inside each segment; (E) marked positions become tokens.
    int n_tok = 0, carry = 0;                  // (the same in every thread: computed from shared memory behind a barrier)
    for (int base = 0; base < n; base += DF_CH) {
        const int L = n - base < DF_CH ? n - base : DF_CH;
        x.par1([&](int t) {                                                                         // (A)
            const int i = base + t;
            int best = 0, dist = 0;
            uint32_t hh = 0xffff;
            if (t < L) {
                const int lim = n - i < DF_MAX_MATCH ? n - i : DF_MAX_MATCH;
                if (lim >= DF_MIN_MATCH) {
                    const uint32_t four = df_word(w32, i);
                    const uint32_t h = (four * 2654435761u) >> (32 - DF_HASH_BITS);
                    hh = h;
                    if (i > 0) { const int l = df_run_len(S.eq, i, lim); if (l >= DF_MIN_MATCH) { best = l; dist = 1; } }
                    const int c = (int)(S.head[h >> 1] >> ((h & 1) << 4) & 0xffffu) - 1;
                    if (best < DF_GOOD_RUN && best < lim && c >= 0 && i - c <= DF_MAX_DIST && df_word(w32, c) == four) {   /* (a long run is taken as it is; most table hits are other 4-grams) */ const int l = df_match_len(w32, c, i, lim < DF_MAX_HASH_MATCH ? lim : DF_MAX_HASH_MATCH); if (l >= DF_MIN_MATCH && l > best) { best = l; dist = i - c; } }
                }
            }
            S.hash[t] = (uint16_t)hh;
            S.mlen[t] = (uint16_t)best; S.mdist[t] = (uint16_t)dist;
            S.jl[0][t] = (uint16_t)(t + (best ? best : 1));
        });
        x.tick(1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 5; ++k)
            x.wpar1([&](int t) {                                                                    // (B)
                // this chunk enters the table; where the next position has the same hash (inside a run) it is the one that counts
                if (k == 0 && S.hash[t] != 0xffff && (t == DF_CH - 1 || S.hash[t + 1] != S.hash[t])) x.atomic_max16(S.head, S.hash[t], (uint32_t)(base + t + 1));
                const int end = (t | 31) + 1 < L ? (t | 31) + 1 : L;     // the end of t's segment (or of the input)
                const int u = S.jl[k][t];
                S.jl[k + 1][t] = u < end ? S.jl[k][u] : (uint16_t)u;
            });
        x.sync();
        x.tick(2);
        {                                                                                           // (C) every thread the same values
            int e = carry - base;
            const int e0 = e;
            for (int g = 0; g < DF_CH / 32; ++g) {
                const int end = 32 * (g + 1) < L ? 32 * (g + 1) : L;
                S.entry[g] = e;
                if (e >= 32 * g && e < end) e = S.jl[5][e];
            }
            if (e0 < L) carry = base + e;
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 6; ++k)
            x.wpar1([&](int t) {                                                                    // (D)
                const int g = t >> 5, end = 32 * (g + 1) < L ? 32 * (g + 1) : L;
                if (k == 0) {
                    if ((t & 31) == 0) { const int e = S.entry[g]; S.markw[g] = e >= 32 * g && e < end ? 1u << (e & 31) : 0u; }
                    return;
                }
                const int u = S.jl[k - 1][t];
                if (((S.markw[g] >> (t & 31)) & 1) && u < end) x.atomic_or(&S.markw[g], 1u << (u & 31));
            });
        x.sync();
        x.tick(3);
        uint32_t n_chunk = 0;
        for (int w = 0; w < DF_CH / 32; ++w) { S.pre[w] = n_chunk; n_chunk += (uint32_t)df_popc(S.markw[w]); }
        x.par1([&](int t) {                                                                         // (E)
            const uint32_t mw = S.markw[t >> 5];
            if (!((mw >> (t & 31)) & 1)) return;
            const uint32_t at = (uint32_t)n_tok + S.pre[t >> 5] + (uint32_t)df_popc(mw & ((1u << (t & 31)) - 1));
            if (S.mlen[t]) {
                tok[at] = 1u << 31 | (uint32_t)(S.mlen[t] - 3) << 16 | (uint32_t)(S.mdist[t] - 1);
                int sym, nx, xv;
                df_len_code(S.mlen[t], sym, nx, xv);
                x.atomic_add(&S.hist[sym], 1u);
                df_dist_code(S.mdist[t], sym, nx, xv);
                x.atomic_add(&S.hist[288 + sym], 1u);
            } else {
                const uint32_t b = df_byte(w32, base + t);
                tok[at] = b;
                x.atomic_add(&S.hist[b], 1u);
            }
        });
        n_tok += (int)n_chunk;
        x.tick(4);
    }
    // ---- 3. codes ----
    x.par(1, [&](int) {
        S.hist[256] += 1;                                        // end of block
        // at least two codes per tree, so that both are complete prefix codes (what every inflater accepts)
        int nd = 0;
        for (int s = 0; s < DF_ND; ++s) nd += S.hist[288 + s] != 0;
        for (int s = 0; nd < 2 && s < DF_ND; ++s) if (!S.hist[288 + s]) { S.hist[288 + s] = 1; ++nd; }
        int nl = 0;
        for (int s = 0; s < DF_NLL; ++s) nl += S.hist[s] != 0;
        for (int s = 0; nl < 2 && s < DF_NLL; ++s) if (!S.hist[s]) { S.hist[s] = 1; ++nl; }
        S.n_used[0] = nl; S.n_used[1] = nd;
    });
    x.par(320, [&](int t) {                                      // rank sort: keys are distinct (the symbol is part of the key)
        const int lo = t < 288 ? 0 : 288, hi = t < 288 ? DF_NLL : 288 + DF_ND;
        if (t >= hi || !S.hist[t]) return;
        const uint32_t key = S.hist[t] << 9 | (uint32_t)(t - lo);
        int r = 0;
        for (int u = lo; u < hi; ++u) r += S.hist[u] && (S.hist[u] << 9 | (uint32_t)(u - lo)) < key;
        S.sorted[lo + r] = key;
    });
    x.tick(5);
    // the two big trees: the merge by one thread each, everything around it across the block
    x.par(320, [&](int t) {
        S.len[t] = 0; S.code[t] = 0;
        if (t < 32) { S.tscr[0][t] = 0; S.tscr[1][t] = 0; }
    });
    x.par(64, [&](int t) {
        if (t & 31) return;
        const int w = t >> 5, o = w ? 288 : 0;
        if (S.n_used[w] >= 2) df_tree_merge(S.sorted + o, S.n_used[w], S.wi + o, S.par_leaf + o, S.par_int + o, S.depth_int + o);
    });
    x.par(320, [&](int t) {
        const int w = t >= 288, o = w ? 288 : 0;
        if (t - o < S.n_used[w]) x.atomic_add(&S.tscr[w][df_tree_leaf_len(S.n_used[w], t - o, 15, S.par_leaf + o, S.depth_int + o)], 1u);
    });
    x.par(64, [&](int t) { if (!(t & 31)) df_tree_limit(S.tscr[t >> 5], 15); });
    x.par(320, [&](int t) {
        const int w = t >= 288, o = w ? 288 : 0;
        if (t - o < S.n_used[w]) S.len[o + (S.sorted[t] & 511)] = (uint8_t)df_tree_len_of_rank(S.tscr[w], 15, (uint32_t)(t - o));
    });
    x.par(320, [&](int t) {
        const int w = t >= 288, o = w ? 288 : 0;
        if (t - o < (w ? DF_ND : DF_NLL)) S.code[t] = (uint16_t)df_tree_code(S.tscr[w], S.len + o, t - o);
    });
    x.tick(6);
    // The block header. The code lengths of both trees, back to back, are run-length coded with the symbols 16 (repeat the
    // previous length 3-6 times), 17 (3-10 zeros), 18 (11-138 zeros): one thread per RUN of equal lengths emits that run's
    // symbols (count, scan, write -- like the tokens), one thread builds the 19-symbol code of those symbols, then every symbol is
    // sized, a scan places it and its bits are ORed into the header words.
    x.par(320, [&](int t) {
        if (t == 0) { int h = DF_NLL; while (h > 257 && !S.len[h - 1]) --h; S.hlit = h; }
        if (t == 32) { int h = DF_ND; while (h > 1 && !S.len[288 + h - 1]) --h; S.hdist = h; }
        if (t >= 64 && t < 64 + DF_NCL) S.clf[t - 64] = 0;
        if (t >= 96 && t < 96 + 80) S.hdr[t - 96] = 0;
    });
    const int hlit = S.hlit, hdist = S.hdist, ns = hlit + hdist;
    auto seq = [&](int p) { return (int)S.len[p < hlit ? p : 288 + p - hlit]; };   // the two length tables back to back
    // the symbols of one run of c equal lengths v, in order, through emit(symbol | extra << 8)
    auto run_symbols = [&](int v, int c, auto emit) {
        if (v == 0) {
            while (c >= 11) { const int r = c < 138 ? c : 138; emit((uint32_t)(18 | (r - 11) << 8)); c -= r; }
            if (c >= 3) { emit((uint32_t)(17 | (c - 3) << 8)); c = 0; }
            for (; c > 0; --c) emit(0u);
        } else {
            emit((uint32_t)v); --c;
            while (c >= 3) { const int r = c < 6 ? c : 6; emit((uint32_t)(16 | (r - 3) << 8)); c -= r; }
            for (; c > 0; --c) emit((uint32_t)v);
        }
    };
    // exclusive prefix sums of S.rcnt[0, 320) into S.roff, S.rgrp[10] = the total (two levels of 32)
    auto scan320 = [&]() {
        x.par(10, [&](int g) { uint32_t sum = 0; for (int k = 32 * g; k < 32 * g + 32; ++k) sum += S.rcnt[k]; S.rgrp[g] = sum; });
        x.par(1, [&](int) { uint32_t c = 0; for (int g = 0; g <= 10; ++g) { const uint32_t v = g < 10 ? S.rgrp[g] : 0u; S.rgrp[g] = c; c += v; } });
        x.par(320, [&](int p) { uint32_t o = S.rgrp[p >> 5]; for (int k = p & ~31; k < p; ++k) o += S.rcnt[k]; S.roff[p] = (uint16_t)o; });
    };
    x.par(320, [&](int p) {
        uint32_t cnt = 0;
        if (p < ns) {
            const int v = seq(p);
            if (p == 0 || seq(p - 1) != v) {
                int c = 1;
                while (p + c < ns && seq(p + c) == v) ++c;
                run_symbols(v, c, [&](uint32_t) { ++cnt; });
            }
        }
        S.rcnt[p] = (uint16_t)cnt;
    });
    scan320();
    x.par(320, [&](int p) {
        if (p >= ns || !S.rcnt[p]) return;
        const int v = seq(p);
        int c = 1;
        while (p + c < ns && seq(p + c) == v) ++c;
        uint32_t at = S.roff[p];
        run_symbols(v, c, [&](uint32_t sy) { S.rle[at++] = (uint16_t)sy; x.atomic_add(&S.clf[sy & 0xff], 1u); });
    });
    const int nr = (int)S.rgrp[10];
    x.par(1, [&](int) {
        // the code of the code lengths: at most 7 bits, 19 symbols -- sorted by insertion (the big trees' scratch is free again)
        int m = 0;
        for (int k = 0; k < DF_NCL; ++k) {
            if (!S.clf[k]) continue;
            const uint32_t key = S.clf[k] << 9 | (uint32_t)k;
            int q = m++;
            while (q > 0 && S.cl_sorted[q - 1] > key) { S.cl_sorted[q] = S.cl_sorted[q - 1]; --q; }
            S.cl_sorted[q] = key;
        }
        df_tree(S.cl_sorted, m, 7, S.cl_len, S.cl_code, DF_NCL, S.wi, S.par_leaf, S.par_int, S.depth_int, S.tscr[2]);
        int hclen = DF_NCL;
        while (hclen > 4 && !S.cl_len[S.order[hclen - 1]]) --hclen;
        DfBits B = {S.hdr, 0, 0, 0};
        B.put(1, 1); B.put(2, 2);                                // BFINAL, BTYPE = dynamic
        B.put((uint32_t)(hlit - 257), 5); B.put((uint32_t)(hdist - 1), 5); B.put((uint32_t)(hclen - 4), 4);
        for (int k = 0; k < hclen; ++k) B.put(S.cl_len[S.order[k]], 3);
        B.finish();                                              // (at most 74 bits: the words behind them are still zero)
        S.hdr_bits = B.n;
    });
    x.par(320, [&](int k) {
        uint32_t nb = 0;
        if (k < nr) { const int sy = S.rle[k] & 0xff; nb = (uint32_t)S.cl_len[sy] + (sy == 16 ? 2u : sy == 17 ? 3u : sy == 18 ? 7u : 0u); }
        S.rcnt[k] = (uint16_t)nb;
    });
    scan320();
    const int fixed_bits = S.hdr_bits;
    x.par(320, [&](int k) {
        if (k >= nr) return;
        const int sy = S.rle[k] & 0xff, xv = S.rle[k] >> 8;
        const uint32_t v = S.cl_code[sy] | (uint32_t)xv << S.cl_len[sy];     // the code and its extra bits: at most 7 + 7 bits
        const uint32_t at = (uint32_t)fixed_bits + S.roff[k];
        x.atomic_or(&S.hdr[at >> 5], v << (at & 31));
        if ((at & 31) + S.rcnt[k] > 32) x.atomic_or(&S.hdr[(at >> 5) + 1], v >> (32 - (at & 31)));
    });
    x.par(1, [&](int) { S.hdr_bits = fixed_bits + (int)S.rgrp[10]; });
    x.tick(7);
    // ---- 4. bits ----
    const int per = (n_tok + DF_CH - 1) / DF_CH;
    x.par(DF_CH, [&](int t) {
        uint32_t b = 0;
        const int lo = t * per, hi = lo + per < n_tok ? lo + per : n_tok;
        for (int k = lo; k < hi; ++k) b += (uint32_t)df_token_bits(S, tok[k]);
        S.part[t] = b;
    });
    x.par(1, [&](int) {
        uint32_t c = 16 + (uint32_t)S.hdr_bits;                  // bit 16 of out32: where the deflate stream starts
        for (int t = 0; t < DF_CH; ++t) { const uint32_t b = S.part[t]; S.part[t] = c; c += b; }
        S.part[DF_CH] = c;
        S.total_bits = c + S.len[256] - 16;
        S.stored = (S.total_bits + 7) / 8 >= (uint32_t)n + 5;
    });
    x.tick(8);
    uint32_t clen;
    if (!S.stored) {
        clen = (S.total_bits + 7) / 8;
        const int n_words = (int)((16 + S.total_bits + 31) / 32);
        x.par(n_words, [&](int w) { out32[w] = 0; });
        const int hdr_words = (S.hdr_bits + 31) / 32;
        x.par(hdr_words > DF_CH ? hdr_words : DF_CH, [&](int t) {
            if (t < hdr_words) {                                 // the header, shifted by the 16 bits
                x.atomic_or(&out32[t], S.hdr[t] << 16);
                if (S.hdr[t] >> 16) x.atomic_or(&out32[t + 1], S.hdr[t] >> 16);
            }
            if (t >= DF_CH) return;
            const int lo = t * per, hi = lo + per < n_tok ? lo + per : n_tok;
            uint32_t bit = S.part[t];
            uint32_t word = bit >> 5;
            int fill = (int)(bit & 31);
            uint64_t acc = 0;
            auto put = [&](uint32_t v, int nb) {
                acc |= (uint64_t)v << fill; fill += nb;
                if (fill >= 32) { x.atomic_or(&out32[word++], (uint32_t)acc); acc >>= 32; fill -= 32; }
            };
            for (int k = lo; k < hi; ++k) {
                const uint32_t tk = tok[k];
                if (!(tk >> 31)) { put(S.code[tk], S.len[tk]); continue; }
                int sym, nx, xv;
                df_len_code((int)((tk >> 16) & 0xff) + 3, sym, nx, xv);
                put(S.code[sym] | (uint32_t)xv << S.len[sym], S.len[sym] + nx);
                df_dist_code((int)(tk & 0xffff) + 1, sym, nx, xv);
                put(S.code[288 + sym] | (uint32_t)xv << S.len[288 + sym], S.len[288 + sym] + nx);
            }
            if (t == DF_CH - 1) put(S.code[256], S.len[256]);    // end of block (the last thread's run ends the stream)
            if (fill) x.atomic_or(&out32[word], (uint32_t)acc);
        });
    } else {
        clen = (uint32_t)n + 5;
        x.par(n, [&](int i) { out[18 + 5 + i] = (uint8_t)df_byte(w32, i); });
        x.par(1, [&](int) {
            uint8_t *d = out + 18;
            d[0] = 1; d[1] = (uint8_t)n; d[2] = (uint8_t)(n >> 8); d[3] = (uint8_t)~d[1]; d[4] = (uint8_t)~d[2];
        });
    }
    x.tick(9);
    // CRC-32 of the input: every thread its slice, then crc(A || B) = crc(A) * x^(8 |B|) + crc(B)
    const int slice = (n + DF_CH - 1) / DF_CH;
    x.par(DF_CH, [&](int t) {
        const int lo = t * slice, hi = lo + slice < n ? lo + slice : n;
        if (lo >= hi) return;
        uint32_t c = 0xffffffffu;
        for (int k = lo; k < hi; ++k) c = S.crc_tab[(c ^ df_byte(w32, k)) & 0xff] ^ (c >> 8);
        c ^= 0xffffffffu;
        if (hi < n) c = df_mulmod(df_x8n(S.x2n, (uint32_t)(n - hi)), c);
        x.atomic_xor(&S.crc, c);
    });
    x.tick(10);
    const uint32_t total = 18 + clen + 8;
    x.par(1, [&](int) {
        const uint8_t H[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
        for (int k = 0; k < 16; ++k) out[k] = H[k];
        out[16] = (uint8_t)(total - 1); out[17] = (uint8_t)((total - 1) >> 8);
        uint8_t *tl = out + 18 + clen;
        const uint32_t c = S.crc, l = (uint32_t)n;
        tl[0] = (uint8_t)c; tl[1] = (uint8_t)(c >> 8); tl[2] = (uint8_t)(c >> 16); tl[3] = (uint8_t)(c >> 24);
        tl[4] = (uint8_t)l; tl[5] = (uint8_t)(l >> 8); tl[6] = (uint8_t)(l >> 16); tl[7] = (uint8_t)(l >> 24);
    });
    x.tick(11);
    return total;
}

} // namespace bsb
