// bsb_cuda.cu -- the sm_100a kernels and the per-batch device pipeline (the product hot path).
//
// One CudaAligner owns one GPU: the index stays resident in HBM for the life of the object, every
// batch flows  H2D -> K1 convert -> K2 seed -> scan -> K3 SA -> K4 chain -> K5 extend
//              [-> pair candidates -> host insert-size stats -> tables H2D] -> K6/7/8 finalise -> D2H.
// There is no CPU execution path in this file: if CUDA is unavailable construction throws.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (no contraction: MAPQ/pairing
// doubles must round exactly like the reference's x86-64 SSE2 arithmetic).
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include "bsb_stages.h"
#include "bsb_warp.cuh"
#include "bsb_extlane.h"
#include "bsb_rescue.h"
#include "bsb_bam.h"
#include "bsb_deflate.h"
#include "bsb_cuda.h"

namespace bsb {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("[E::bsbolt_b200] CUDA error: ") + cudaGetErrorString(e_) + " (" #x ") at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    void ensure(size_t n)
    {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        size_t c = n + n / 4 + 64;
        CK(cudaMalloc((void **)&p, c * sizeof(T)));
        cap = c;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { release(); }
};

// ------------------------------------------------------------------------------------------------
// kernels: thin launch shells around the stage bodies of bsb_stages.h
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int owner_of(const uint32_t *off, int n, uint32_t g)
{   // largest r with off[r] <= g   (off is non-decreasing, off[n] > g)
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// K1: sixteen bases per thread (one 128-bit load, two 128-bit stores); the read that owns the first base is found once
// and the walk steps into the next read(s) where the chunk crosses a boundary
__global__ void k_convert(BatchDev B, uint32_t n_bases)
{
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    if (i0 >= n_bases) return;
    int r = owner_of(B.seq_off, B.n, i0);
    uint32_t next = B.seq_off[r + 1];
    int pat = B.pattern[r];
    const uint32_t m = n_bases - i0 < 16u ? n_bases - i0 : 16u;
    uint32_t in[4];
    if (m == 16u) { const uint4 v = *reinterpret_cast<const uint4 *>(B.bases + i0); in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w; }
    else { in[0] = in[1] = in[2] = in[3] = 0; for (uint32_t t = 0; t < m; ++t) in[t >> 2] |= (uint32_t)(uint8_t)B.bases[i0 + t] << ((t & 3) << 3); }
    uint32_t so[4] = {0, 0, 0, 0}, oo[4] = {0, 0, 0, 0};
#pragma unroll
    for (uint32_t t = 0; t < 16u; ++t) {
        if (t < m) {
            const uint32_t i = i0 + t;
            while (i >= next) { ++r; next = B.seq_off[r + 1]; pat = B.pattern[r]; }
            unsigned char c = (unsigned char)(in[t >> 2] >> ((t & 3) << 3));
            oo[t >> 2] |= (uint32_t)nt4_code(c) << ((t & 3) << 3);
            if (pat) { if (c == 'G') c = 'A'; }
            else { if (c == 'C') c = 'T'; }
            const uint8_t code = nt4_code(c);
            so[t >> 2] |= (uint32_t)(code == 5 ? 4 : code) << ((t & 3) << 3);   // as stage_convert_base (bsb_stages.h)
        }
    }
    if (m == 16u) {
        *reinterpret_cast<uint4 *>(B.seq + i0) = make_uint4(so[0], so[1], so[2], so[3]);
        *reinterpret_cast<uint4 *>(B.oseq + i0) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    } else
        for (uint32_t t = 0; t < m; ++t) { B.seq[i0 + t] = (uint8_t)(so[t >> 2] >> ((t & 3) << 3)); B.oseq[i0 + t] = (uint8_t)(oo[t >> 2] >> ((t & 3) << 3)); }
}

// ---- K2, two-item form (bsb_seed3.h) ---------------------------------------------------------
// 4-bit packed copy of the converted reads, 16 bases per 64-bit word; read r starts at word (seq_off[r] >> 4) + r
// (a closed form that never overlaps: ceil(len/16) <= (len >> 4) + 1). A lane keeps one word in registers.
__global__ void k_pack4(BatchDev B, uint64_t *seq4, uint32_t n_words)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    int lo = 0, hi = B.n;                       // largest r with start(r) <= w
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((B.seq_off[mid] >> 4) + (uint32_t)mid <= w) lo = mid; else hi = mid;
    }
    const uint32_t beg = B.seq_off[lo], len = B.seq_off[lo + 1] - beg;
    const uint32_t k = w - ((beg >> 4) + (uint32_t)lo);
    if (k * 16 >= len) return;                  // padding word between two reads
    uint64_t v = 0;
    const uint32_t m = min(16u, len - k * 16);
    for (uint32_t t = 0; t < m; ++t) v |= (uint64_t)(B.seq[beg + k * 16 + t] & 15) << (t << 2);
    seq4[w] = v;
}

// One 64-bit word (16 bases) of the read in registers, the word the sweep will need next already in flight:
// a lane that stalls on a load inside the divergent part of the machine stalls the other 31 lanes with it.
struct BasesPacked {
    const uint64_t *w; uint64_t cur, nxt; int widx, nidx, nwords;
    __device__ __forceinline__ void open(const uint64_t *words, int len) { w = words; nwords = (len + 15) >> 4; widx = -1; nidx = -1; cur = nxt = 0; }
    __device__ __forceinline__ int get(int i)
    {
        const int wi = i >> 4;
        if (wi != widx) {
            cur = wi == nidx ? nxt : __ldg(w + wi);
            nidx = wi + (wi < widx ? -1 : 1);
            widx = wi;
            if (nidx >= 0 && nidx < nwords) nxt = __ldg(w + nidx); else nidx = -1;
        }
        return (int)(cur >> ((i & 15) << 2)) & 15;
    }
};

// Interval list of one lane (bsb_seed3.h). Entries are 12 bytes (40-bit coordinates). Shared memory holds a ring of the
// SCAP most recently pushed ranks, lane-interleaved so that lanes at the same rank hit different banks; a push beyond
// SCAP first moves the rank it overwrites to the lane's HBM spill row. The backward sweep lives at the top of the
// list (ranks nf-1 downwards, shrinking row by row), i.e. in the ring; only its first rows reach into the spill row.
constexpr int SEED3_BLOCK = 64;
template <int SCAP>
struct ListRing {
    uint32_t *s; uint4 *g; int total;
    __device__ __forceinline__ int cap() const { return total; }
    template <class U> __device__ __forceinline__ static uint32_t pack(U x0, U x2, int end)
    {
        uint32_t m = (uint32_t)end << 16;
        if (sizeof(U) == 8) m |= ((uint32_t)((uint64_t)x0 >> 32) & 0xff) | ((uint32_t)((uint64_t)x2 >> 32) & 0xff) << 8;
        return m;
    }
    template <class U> __device__ __forceinline__ static void unpack(uint32_t a, uint32_t b, uint32_t m, U &x0, U &x2, int &end)
    {
        if (sizeof(U) == 8) { x0 = (U)((uint64_t)(m & 0xff) << 32 | a); x2 = (U)((uint64_t)((m >> 8) & 0xff) << 32 | b); }
        else { x0 = (U)a; x2 = (U)b; }
        end = (int)(m >> 16);
    }
    template <class U> __device__ __forceinline__ void push(int p, U x0, U x2, int end)
    {
        uint32_t *e = s + (p & (SCAP - 1)) * 3 * SEED3_BLOCK;
        if (p >= SCAP) g[p - SCAP] = make_uint4(e[0], e[SEED3_BLOCK], e[2 * SEED3_BLOCK], 0);
        e[0] = (uint32_t)x0; e[SEED3_BLOCK] = (uint32_t)x2; e[2 * SEED3_BLOCK] = pack(x0, x2, end);
    }
    template <class U> __device__ __forceinline__ void get(int p, int nf, U &x0, U &x2, int &end) const
    {
        if (p + SCAP >= nf) {
            const uint32_t *e = s + (p & (SCAP - 1)) * 3 * SEED3_BLOCK;
            unpack(e[0], e[SEED3_BLOCK], e[2 * SEED3_BLOCK], x0, x2, end);
        } else {
            const uint4 v = g[p];
            unpack(v.x, v.y, v.z, x0, x2, end);
        }
    }
    template <class U> __device__ __forceinline__ void set(int p, int nf, U x0, U x2, int end)
    {
        if (p + SCAP >= nf) {
            uint32_t *e = s + (p & (SCAP - 1)) * 3 * SEED3_BLOCK;
            e[0] = (uint32_t)x0; e[SEED3_BLOCK] = (uint32_t)x2; e[2 * SEED3_BLOCK] = pack(x0, x2, end);
        } else g[p] = make_uint4((uint32_t)x0, (uint32_t)x2, pack(x0, x2, end), 0);
    }
};

template <bool COMPACT> struct SeedCoord { typedef uint64_t type; };
template <> struct SeedCoord<true> { typedef uint32_t type; };

// Every lane runs one work item (item < n: passes 1+2 of read `item`; item >= n: pass 3 of read item - n). The warp
// draws items 32 at a time from a global counter and hands them to its lanes as they finish; all lanes meet at the
// single extension site. COMPACT: sector-sized occ blocks and 32-bit coordinates (index < 2^32 symbols).
template <int SCAP, int MINB, bool COMPACT>
__global__ void __launch_bounds__(SEED3_BLOCK, MINB) k_seed3(Opt opt, IndexView ix, BatchDev B, const uint64_t *seq4, uint4 *spill, int ltotal,
                                                              int *next_item, int32_t *cnt_a, int32_t *cnt_b, unsigned long long *work)
{
    extern __shared__ uint32_t sm_list[];
    uint32_t n_ext = 0, n_two = 0, n_two_ref = 0;  // FM extensions of this lane; those whose two ranks lie in different occ blocks (this layout / the reference's 128-symbol blocks)
    typedef typename SeedCoord<COMPACT>::type U;
    typedef Seeder3<BasesPacked, ListRing<SCAP>, U> Machine;
    Machine sm;
    sm.L.s = sm_list + threadIdx.x; sm.L.total = ltotal;
    sm.L.g = spill + (size_t)(blockIdx.x * SEED3_BLOCK + threadIdx.x) * (size_t)ltotal;
    sm.st = Machine::DONE; sm.err = 0; sm.n_out = 0;
    const unsigned lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    const int n_items = 2 * B.n;
    int item = -1, pool_next = 0, pool_end = 0;   // the warp's drawn items [pool_next, pool_end): same values in all lanes
    bool dry = false;                              // the global counter has run out
    bool need = false;                             // this lane has an extension pending
    for (;;) {
        // ---- converged: close finished items, hand out new ones ----
        const bool idle = !need;
        if (idle && item >= 0) {
            const int r = item < B.n ? item : item - B.n;
            (item < B.n ? cnt_a : cnt_b)[r] = sm.n_out;
            if (sm.err) B.err[r] = sm.err;
            item = -1;
        }
        unsigned want = __ballot_sync(0xffffffffu, idle);
        while (want) {
            if (pool_next == pool_end) {
                if (dry) break;
                int base = 0;
                if (lane == 0) base = atomicAdd(next_item, 32);
                base = __shfl_sync(0xffffffffu, base, 0);
                pool_next = min(base, n_items); pool_end = min(base + 32, n_items);
                if (pool_next == pool_end) { dry = true; break; }
            }
            const int rank = __popc(want & lt_mask);
            const bool take = idle && item < 0 && pool_next + rank < pool_end;
            if (take) {
                item = pool_next + rank;
                sm.n_out = 0; sm.err = 0;
                const int r = item < B.n ? item : item - B.n;
                const uint32_t beg = B.seq_off[r];
                const int len = (int)(B.seq_off[r + 1] - beg);
                if (len >= opt.min_seed_len) {
                    sm.q.open(seq4 + (beg >> 4) + r, len);
                    sm.init(opt, len, B.intv + (size_t)r * B.intv_cap, B.intv_cap, item >= B.n);
                    need = sm.advance(opt, ix);
                }   // else: nothing to seed; the lane closes the empty item and draws again next round
            }
            const unsigned took = __ballot_sync(0xffffffffu, take);
            pool_next += __popc(took);
            want &= ~took;
        }
        const unsigned busy = __ballot_sync(0xffffffffu, need || item >= 0);
        if (!busy) break;
        // ---- the one extension site, then every lane runs its machine up to its next extension ----
        if (need) {
            U xa, xb, s, na, nb, sz;
            sm.request(xa, xb, s);
            ++n_ext;
            n_two += COMPACT ? ((uint32_t)(xa - 1) >> 6) != ((uint32_t)(xa - 1 + s) >> 6) : ((uint64_t)(xa - 1) >> 7) != ((uint64_t)(xa - 1 + s) >> 7);
            n_two_ref += ((uint64_t)(xa - 1) >> 7) != ((uint64_t)(xa - 1 + s) >> 7);
            fm_extend_any(ix, xa, xb, s, sm.c, na, nb, sz);
            need = sm.step(opt, ix, na, nb, sz);
        }
        __syncwarp();
    }
    // algorithmic work of the launch, counted: extensions and occ-block sectors (measurement; two adds per extension)
    n_ext = __reduce_add_sync(0xffffffffu, n_ext); n_two = __reduce_add_sync(0xffffffffu, n_two); n_two_ref = __reduce_add_sync(0xffffffffu, n_two_ref);
    if (lane == 0 && work) { atomicAdd(work, (unsigned long long)n_ext); atomicAdd(work + 1, (unsigned long long)n_two); atomicAdd(work + 3, (unsigned long long)n_two_ref); }
}

// index-load time: the sector-sized occ blocks (bsb_index.h) from the reference-layout BWT
__global__ void k_make_occ32(const uint32_t *bwt, uint64_t n_blocks32, uint32_t *occ32)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks32) return;
    uint32_t o[8];
    occ32_make_block(bwt, b, o);
    uint4 *dst = reinterpret_cast<uint4 *>(occ32 + b * 8);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]); dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// One warp per read: merges the two parts of the read's interval list (front: item A, back: item B), sorts it by
// `info` (rank by counting; up to four entries per lane, keys staged in shared memory) and runs the tail of the stage
// (bwamem.c:269-283): seed count, and l_rep = length of the union of the repetitive intervals (a bitmap per warp).
constexpr int FIN_MAX = 128, FIN_WORDS = 24;   // entries handled in registers; bitmap words (reads <= 768 bp)
__global__ void __launch_bounds__(128) k_seed3_finish(Opt opt, BatchDev B, const int32_t *cnt_a, const int32_t *cnt_b)
{
    __shared__ uint64_t s_keys[4][FIN_MAX];
    __shared__ uint32_t s_cov[4][FIN_WORDS];
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, wib = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= B.n) return;
    if (B.err[r]) return;
    Intv *mem = B.intv + (size_t)r * B.intv_cap;
    const int na = cnt_a[r], nb = cnt_b[r], n = na + nb;
    if (na + 2 * nb > B.intv_cap) { if (lane == 0) B.err[r] = ERR_INTV_OVERFLOW; return; }
    const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    if (n > FIN_MAX || len > FIN_WORDS * 32) {   // very long lists / reads: the serial form
        if (lane == 0) {
            const int m = seed3_merge_sort(mem, B.intv_cap, na, nb);
            seed_finish(opt, B, r, mem, m, 0);
        }
        return;
    }
    uint64_t *keys = s_keys[wib];
    uint32_t *cov = s_cov[wib];
    Intv v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = lane + 32 * q;
        if (e < n) { v[q] = mem[e < na ? e : B.intv_cap - 1 - (e - na)]; keys[e] = v[q].info; }
    }
    if (lane < FIN_WORDS) cov[lane] = 0;
    __syncwarp();
    int rank[4] = {0, 0, 0, 0};
    for (int f = 0; f < n; ++f) {
        const uint64_t o = keys[f];
#pragma unroll
        for (int q = 0; q < 4; ++q) rank[q] += (o < v[q].info) || (o == v[q].info && f < lane + 32 * q);
    }
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (lane + 32 * q < n) {
            mem[rank[q]] = v[q];
            const uint64_t occ = v[q].x2;
            if (occ > (uint64_t)opt.max_occ) {   // repetitive interval: at most max_occ of its hits, its span counts towards l_rep
                const int64_t step = (int64_t)(occ / opt.max_occ);
                const int64_t c = ((int64_t)occ + step - 1) / step;
                cnt += (int)(c < opt.max_occ ? c : opt.max_occ);
                const int sb = (int)(v[q].info >> 32), se = (int)(uint32_t)v[q].info;
                for (int w = sb >> 5; w <= (se - 1) >> 5 && se > sb; ++w) {
                    const int lo = max(sb - 32 * w, 0), hi = min(se - 32 * w, 32);
                    atomicOr(&cov[w], (hi >= 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u));
                }
            } else cnt += (int)occ;
        }
    }
    __syncwarp();
    int l_rep = lane < FIN_WORDS ? __popc(cov[lane]) : 0;
    for (int o = 16; o; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); l_rep += __shfl_xor_sync(0xffffffffu, l_rep, o); }
    if (lane == 0) { B.n_intv[r] = n; B.l_rep[r] = l_rep; B.n_seed[r] = cnt; }
}

// K3, one thread per seed. The read that owns the block's first seed is found once per block; the seeds of a block span a
// handful of reads, so every thread finishes its own search inside a short window of seed_off kept in shared memory (and
// falls back to the full search beyond it).
__global__ void __launch_bounds__(128) k_sa(Opt opt, IndexView ix, BatchDev B, uint32_t n_seeds)
{
    constexpr int WIN = 64;
    __shared__ uint32_t s_off[WIN + 1];
    __shared__ int s_r0;
    const uint32_t g0 = blockIdx.x * blockDim.x, g = g0 + threadIdx.x;
    if (threadIdx.x == 0) s_r0 = owner_of(B.seed_off, B.n, g0);
    __syncthreads();
    const int r0 = s_r0;
    if (threadIdx.x <= WIN) s_off[threadIdx.x] = r0 + (int)threadIdx.x <= B.n ? B.seed_off[r0 + threadIdx.x] : 0xffffffffu;
    __syncthreads();
    if (g >= n_seeds) return;
    int r;
    if (s_off[WIN] > g) {            // largest k with s_off[k] <= g (s_off[0] <= g0 <= g)
        int lo = 0, hi = WIN;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= g) lo = mid; else hi = mid; }
        r = r0 + lo;
    } else r = owner_of(B.seed_off, B.n, g);
    stage_sa(opt, ix, B, r, g);
}

// Warps of the warp-per-read kernels draw their reads from a global counter (reads differ by orders of magnitude in
// seeds and chains; a fixed stride leaves most warps idle behind the few that met the heavy reads). ctr == nullptr:
// fixed stride.
__device__ __forceinline__ int warp_next_read(int *ctr, int cur, int stride)
{
    if (!ctr) return cur + stride;
    int r = 0;
    if ((threadIdx.x & 31) == 0) r = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, r, 0);
}

// the same for kernels whose lanes each take one item: the warp draws 32 consecutive items at a time
__device__ __forceinline__ unsigned warp_next_group(int *ctr, unsigned cur, unsigned stride)
{
    if (!ctr) return cur + stride;
    unsigned r = 0;
    if ((threadIdx.x & 31) == 0) r = (unsigned)atomicAdd(ctr, 32);
    return __shfl_sync(0xffffffffu, r, 0);
}

// K4 (bsb_warp.cuh): a warp takes a group of 32 consecutive reads through the three phases
__global__ void __launch_bounds__(128, 6) k_chain_warp(Opt opt, IndexView ix, BatchDev B, int *ctr, int32_t *aux)
{
    __shared__ ChainSmem sm[4];
    const int wib = threadIdx.x >> 5;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned base = warp_next_group(ctr, gw * 32 - nw * 32, nw * 32); base < (unsigned)B.n; base = warp_next_group(ctr, base, nw * 32))
        stage_chain_group(opt, ix, B, (int)base, sm[wib], aux);
}



// K5, warp per read: rows of the banded extension across the lanes, (h,e) rows + query in shared memory
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_extend_warp(Opt opt, IndexView ix, BatchDev B, int32_t *eh, int max_q, int smem_per_warp, const int32_t *order,
                                                            int *ctr)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int wib = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    uint8_t *mine = smem + (size_t)wib * smem_per_warp;
    WarpDp S;
    S.H = (int32_t *)mine; S.E = S.H + (max_q + 1); S.qs = (uint8_t *)(S.E + (max_q + 1));
    DpScratch dp = {eh + (size_t)gw * 32 * 2 * (max_q + 1), nullptr, 0, max_q};
    DpScratch dp_lane = {eh + ((size_t)gw * 32 + (threadIdx.x & 31)) * 2 * (max_q + 1), nullptr, 0, max_q};
    for (unsigned base = warp_next_group(ctr, (unsigned)gw * 32 - (unsigned)nw * 32, (unsigned)nw * 32); base < (unsigned)B.n;
         base = warp_next_group(ctr, base, (unsigned)nw * 32))
        stage_extend_group(opt, ix, B, (int)base, min(32, B.n - (int)base), order, S, dp, dp_lane);
}

// K5, lane per read (bsb_extlane.h): every lane runs the extension machine of one read; the warp draws reads 32 at a time from
// a global counter and hands them to its lanes as they finish; between two DP rows a lane runs its own (divergent, short)
// control flow, then all lanes that have a row to fill meet in row(). The (h,e) rows live in shared memory, word
// j * 32 + lane of the warp's tile. Finished reads carry their regions before mem_sort_dedup_patch; k_extend_tail ends them.
constexpr int XL_WARPS = 2;
__global__ void __launch_bounds__(XL_WARPS * 32) k_extend_lanes(Opt opt, IndexView ix, BatchDev B, int max_q, int row_cap, int *ctr, int chunk, int ctl_mask, int ctl_lanes,
                                                                 unsigned long long *work)
{
    extern __shared__ uint32_t xl_rows[];
    typedef ExtLane<PackedRow<32>> Machine;
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lt_mask = (1u << lane) - 1u;
    Machine L;
    L.state = Machine::IDLE;
    L.n_cells = 0;
    L.H.p = xl_rows + (size_t)wib * 32 * (size_t)row_cap + lane;
    L.row_cap = row_cap;
    {   // per-query-length tables behind the row tiles: max_gap, band limits of the left / right extension
        int *tab = reinterpret_cast<int *>(xl_rows + (size_t)XL_WARPS * 32 * (size_t)row_cap);
        const int nt = max_q + 2, amax = ext_amax(opt);
        for (int q = threadIdx.x; q < nt; q += blockDim.x) ext_tables_fill(opt, amax, q, tab, tab + nt, tab + 2 * nt);
        L.T.gap = tab; L.T.wl = tab + nt; L.T.wr = tab + 2 * nt; L.T.n = nt; L.T.amax = amax;
        __syncthreads();
    }
    int pool_next = 0, pool_end = 0;   // the warp's drawn reads [pool_next, pool_end): same values in all lanes
    bool dry = false;
    for (unsigned iter = 0;; ++iter) {
        unsigned rows = __ballot_sync(0xffffffffu, L.state == Machine::ROW);
        const bool more_reads = !(dry && pool_next == pool_end);
        // lanes a control phase would move: those between two rows, and idle ones while reads remain. The phase is divergent,
        // so it waits until several lanes need it (or nothing else can run, or every fourth step at the latest)
        const unsigned can = __ballot_sync(0xffffffffu, L.state != Machine::ROW && (L.state != Machine::IDLE || more_reads));
        if (can && (!rows || __popc(can) >= ctl_lanes || (iter & ctl_mask) == 0)) {
            for (;;) {                   // idle lanes take reads (a read without seeds leaves its lane idle: it draws again)
                const unsigned want = __ballot_sync(0xffffffffu, L.state == Machine::IDLE);
                if (!want) break;
                if (pool_next == pool_end) {
                    if (dry) break;
                    int base = 0;
                    if (lane == 0) base = atomicAdd(ctr, 32);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    pool_next = min(base, B.n); pool_end = min(base + 32, B.n);
                    if (pool_next == pool_end) { dry = true; break; }
                }
                const int rank = __popc(want & lt_mask);
                const bool take = L.state == Machine::IDLE && pool_next + rank < pool_end;
                if (take) L.begin_read(B, pool_next + rank);
                pool_next += __popc(__ballot_sync(0xffffffffu, take));
            }
            if (L.state != Machine::IDLE && L.state != Machine::ROW) L.advance(opt, ix, B, max_q);
            __syncwarp();
            while (__ballot_sync(0xffffffffu, L.in_init())) {   // the lanes that have just set up an extension fill its initial row together
                if (L.in_init()) L.init_step(opt, 32);
                __syncwarp();
            }
            rows = __ballot_sync(0xffffffffu, L.state == Machine::ROW);
        }
        if (!rows) {
            if (dry && pool_next == pool_end && !__ballot_sync(0xffffffffu, L.state != Machine::IDLE)) break;
            continue;
        }
        if (L.state == Machine::ROW) L.step(opt, ix, chunk);
        __syncwarp();
    }
    const unsigned cells = __reduce_add_sync(0xffffffffu, L.n_cells);   // (a warp fills far fewer than 2^32 cells)
    if (lane == 0 && work) atomicAdd(work + 2, (unsigned long long)cells);
}

// mem_sort_dedup_patch + ALT marks, one thread per read (serial by nature: two introsorts, the redundancy scan, now and then a
// score-only global alignment in the thread's own DP scratch)
__global__ void __launch_bounds__(128) k_extend_tail(Opt opt, IndexView ix, BatchDev B, int32_t *eh, int max_q)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, nw = gridDim.x * blockDim.x;
    DpScratch dp = {eh + (size_t)w * 2 * (max_q + 1), nullptr, 0, max_q};
    for (int r = w; r < B.n; r += nw) extend_tail(opt, ix, B, r, dp);
}

// K4b: mem_flt_chained_seeds (bwamem.c:602-619). Launched only when a read of the batch can qualify (reads of about
// 720 bp and more, or -W): a warp takes a read, each lane one of its chains; the flanked 16-bit Smith-Waterman of a seed
// covers fewer than 200 x 200 cells, its four rows live in the lane's local memory.
__global__ void __launch_bounds__(128) k_seed_sw(Opt opt, IndexView ix, BatchDev B, const double *log_tab, int n_log)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = gw; r < B.n; r += nw) {
        const int nc = B.n_chain[r];
        const int had_err = B.err[r];
        __syncwarp();
        if (nc == 0 || had_err) continue;
        const uint32_t so = B.seed_off[r];
        const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
        int err = 0;
        const int min_hsp = seed_sw_min_score(opt, len, log_tab, n_log, &err);
        if (err) { if (lane == 0) B.err[r] = err; continue; }
        if (min_hsp < 0) continue;
        int32_t rows[4 * SEED_SW_CAP];
        SwScratch ws = {rows, rows + SEED_SW_CAP, rows + 2 * SEED_SW_CAP, rows + 3 * SEED_SW_CAP, nullptr, SEED_SW_CAP, 0};
        for (int i = lane; i < nc; i += 32) {
            filter_chained_seeds(opt, ix, len, B.seq + B.seq_off[r], 1, B.chains + so + i, B.cseeds + so, min_hsp, ws, &err);
            if (err) B.err[r] = err;
        }
    }
}

__global__ void k_pestat(Opt opt, IndexView ix, BatchDev B)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < (B.n >> 1)) stage_pestat(opt, ix, B, p);
}

struct FinalLayout {       // byte offsets inside one worker's scratch block (selection / pairing kernel)
    size_t eh, cnt, has_alt, zz, pv, pu, sw, swb, rev, wregs, total;
    int max_q, reg_cap, pair_cap, sw_cap, sw_b, wreg_stride;
};

__device__ __forceinline__ void make_ws(const FinalLayout &L, uint8_t *blk, FinalWS &ws, AlnReg *&wregs)
{
    ws.dp.eh = (int32_t *)(blk + L.eh); ws.dp.z = nullptr; ws.dp.z_cap = 0; ws.dp.max_q = L.max_q;
    ws.cigar = nullptr; ws.cigar_cap = 0; ws.md = nullptr; ws.md_cap = 0; ws.xb = nullptr; ws.xb_cap = 0;
    ws.cnt = (int32_t *)(blk + L.cnt); ws.has_alt = (int8_t *)(blk + L.has_alt); ws.z = (int32_t *)(blk + L.zz);
    ws.pv = (Pair64 *)(blk + L.pv); ws.pu = (Pair64 *)(blk + L.pu); ws.pair_cap = L.pair_cap; ws.reg_cap = L.reg_cap;
    int32_t *sw = (int32_t *)(blk + L.sw);
    ws.sw.H0 = sw; ws.sw.H1 = sw + L.sw_cap; ws.sw.E = sw + 2 * L.sw_cap; ws.sw.Hmax = sw + 3 * L.sw_cap;
    ws.sw.b = (uint64_t *)(blk + L.swb); ws.sw.cap = L.sw_cap; ws.sw.cap_b = L.sw_b;
    ws.rev = blk + L.rev;
    wregs = (AlnReg *)(blk + L.wregs);
}

// K6a: global alignments of the queued alignments that need one (about one in ten), warp-cooperative
__global__ void __launch_bounds__(128) k_tasks_dp(Opt opt, IndexView ix, BatchDev B, unsigned int n_tasks, uint8_t *zbuf, long z_cap, int max_q, int smem_per_warp,
                                                  uint32_t *slots, int slot_cap, int32_t *n_cig, int *ctr)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    uint8_t *mine = smem + (size_t)wib * smem_per_warp;
    WarpTask S;
    S.H = (int32_t *)mine; S.E = S.H + (max_q + 1); S.qs = (uint8_t *)(S.E + (max_q + 1));
    S.cigar = nullptr; S.cigar_cap = 0; S.md = nullptr; S.md_cap = 0; S.xb = nullptr; S.xb_cap = 0;
    uint8_t *z = zbuf + (size_t)gw * z_cap;
    for (unsigned int base = warp_next_group(ctr, gw * 32 - nw * 32, nw * 32); base < n_tasks; base = warp_next_group(ctr, base, nw * 32)) {
        const unsigned int k = base + lane;
        bool dp = false;
        if (k < n_tasks) { const AlnTask t = B.tasks.a[k]; dp = !B.out[t.read].err && !task_is_trivial(opt, t); }
        unsigned m = __ballot_sync(0xffffffffu, dp);
        while (m) {
            const unsigned int kk = base + (unsigned)__ffs(m) - 1;
            m &= m - 1;
            stage_task_dp_warp(opt, ix, B, kk, S, z, z_cap, max_q, slots + (size_t)kk * slot_cap, slot_cap, n_cig + kk);
        }
    }
}

// K6b + K8b: one thread per queued alignment
__global__ void __launch_bounds__(128) k_tasks_finish(Opt opt, IndexView ix, BatchDev B, unsigned int n_tasks, uint32_t *slots, int slot_cap, const int32_t *n_cig,
                                                      char *text, int md_cap, int xb_cap, int *ctr)
{
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x, nw = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    char *md = text + (size_t)w * (size_t)(md_cap + xb_cap);
    for (unsigned int base = warp_next_group(ctr, (w & ~31u) - nw, nw); base < n_tasks; base = warp_next_group(ctr, base, nw)) {
        const unsigned int k = base + lane;
        if (k < n_tasks) stage_task_finish(opt, ix, B, k, slots + (size_t)k * slot_cap, n_cig + k, md, md_cap, md + md_cap, xb_cap);
    }
}

__global__ void __launch_bounds__(32) k_final_se(Opt opt, IndexView ix, BatchDev B, FinalLayout L, uint8_t *scratch)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, nw = gridDim.x * blockDim.x;
    FinalWS ws; AlnReg *wregs;
    make_ws(L, scratch + (size_t)w * L.total, ws, wregs);
    for (int r = w; r < B.n; r += nw) stage_final_se(opt, ix, B, r, ws, wregs);
}

// K7 + K8a, thread per pair. Pairs that need mate-rescue Smith-Waterman are not finalised here: they are queued (heavy,
// n_heavy) for the warp-cooperative kernel below, before anything has been written for them.
template <int MINB>
__global__ void __launch_bounds__(32, MINB) k_final_pe(Opt opt, IndexView ix, BatchDev B, FinalLayout L, uint8_t *scratch, int32_t *heavy, int *n_heavy,
                                                        int *ctr)
{
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x, nw = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    const unsigned np = (unsigned)(B.n >> 1);
    FinalWS ws; AlnReg *wregs;
    make_ws(L, scratch + (size_t)w * L.total, ws, wregs);
    for (unsigned base = warp_next_group(ctr, (w & ~31u) - nw, nw); base < np; base = warp_next_group(ctr, base, nw)) {
        const unsigned p = base + lane;
        if (p < np && stage_final_pe(opt, ix, B, (int)p, ws, wregs, heavy != nullptr)) heavy[atomicAdd(n_heavy, 1)] = (int32_t)p;
    }
}

// K7 for the queued pairs: one warp per pair, rescue Smith-Waterman across the lanes (bsb_warp.cuh)
__global__ void __launch_bounds__(128) k_final_pe_heavy(Opt opt, IndexView ix, BatchDev B, FinalLayout L, uint8_t *scratch, const int32_t *heavy, const int *n_heavy,
                                                        int *ctr)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int wib = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    FinalWS ws; AlnReg *wregs;
    make_ws(L, scratch + (size_t)gw * L.total, ws, wregs);
    WarpSw sw;
    int32_t *base = reinterpret_cast<int32_t *>(smem) + (size_t)wib * (4 * L.sw_cap + (L.sw_cap + 3) / 4);
    sw.H0 = base; sw.H1 = base + L.sw_cap; sw.E = base + 2 * L.sw_cap; sw.Hmax = base + 3 * L.sw_cap;
    sw.qs = reinterpret_cast<uint8_t *>(base + 4 * L.sw_cap); sw.cap = L.sw_cap;
    sw.b = ws.sw.b; sw.cap_b = ws.sw.cap_b;
    const int n = *n_heavy;
    for (int k = warp_next_read(ctr, gw - nw, nw); k < n; k = warp_next_read(ctr, k, nw)) stage_final_pe_heavy(opt, ix, B, heavy[k], ws, wregs, sw);
}

// ---- mate rescue by jobs (bsb_final.h: RescueJob; bsb_rescue.h: the Smith-Waterman machine) --------------------------------
// 1. k_rescue_enum, thread per queued pair, twice: every Smith-Waterman the pair can need is counted, then (after a scan
//    over the counts) written into the pair's own block of the job list;
// 2. k_rescue_sw, lane per job: the jobs are drawn from a counter, each lane fills its own tile of shared memory;
// 3. k_rescue_replay, thread per queued pair: mem_sam_pe with the rescue results looked up instead of computed. A pair with
//    a job the lane kernel could not take (saturated score, overlong sub-optimal list) goes on to the warp-per-pair kernel.
__global__ void __launch_bounds__(32, 16) k_rescue_enum(Opt opt, IndexView ix, BatchDev B, FinalLayout L, uint8_t *scratch, const int32_t *heavy, int n_heavy,
                                                        RescueJob *jobs, const uint32_t *job_off, uint32_t *job_cnt, int *ctr)
{
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x, nw = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    FinalWS ws; AlnReg *wregs;
    make_ws(L, scratch + (size_t)w * L.total, ws, wregs);
    for (unsigned base = warp_next_group(ctr, (w & ~31u) - nw, nw); base < (unsigned)n_heavy; base = warp_next_group(ctr, base, nw)) {
        const unsigned k = base + lane;
        if (k >= (unsigned)n_heavy) continue;
        unsigned int cnt = 0;
        RescueSink sink = {jobs ? jobs + job_off[k] : nullptr, &cnt, jobs ? job_off[k + 1] - job_off[k] : 0u, (int32_t)k, 0};
        stage_final_pe(opt, ix, B, heavy[k], ws, wregs, false, &sink, nullptr);
        if (!jobs) job_cnt[k] = cnt;
    }
}

__global__ void __launch_bounds__(32) k_rescue_sw(Opt opt, IndexView ix, BatchDev B, const RescueJob *jobs, SwResult *res, int n_jobs, int cap_cells,
                                                  uint64_t *blist, int cap_b, int *ctr, int chunk)
{
    extern __shared__ uint32_t rs_tile[];
    typedef SwLane<PackedRow<32>> Machine;
    const unsigned lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    Machine L;
    L.state = Machine::IDLE;
    L.W.p = rs_tile + lane; L.cap_cells = cap_cells;
    L.b = blist + ((size_t)blockIdx.x * 32 + lane) * (size_t)cap_b; L.cap_b = cap_b;
    int mine = -1, pool_next = 0, pool_end = 0;
    bool dry = false;
    for (;;) {
        for (;;) {                       // idle lanes take jobs
            const unsigned want = __ballot_sync(0xffffffffu, L.state == Machine::IDLE);
            if (!want) break;
            if (pool_next == pool_end) {
                if (dry) break;
                int base = 0;
                if (lane == 0) base = atomicAdd(ctr, 32);
                base = __shfl_sync(0xffffffffu, base, 0);
                pool_next = min(base, n_jobs); pool_end = min(base + 32, n_jobs);
                if (pool_next == pool_end) { dry = true; break; }
            }
            const int rank = __popc(want & lt_mask);
            const bool take = L.state == Machine::IDLE && pool_next + rank < pool_end;
            if (take) {
                mine = pool_next + rank;
                const RescueJob jb = jobs[mine];
                L.begin(opt, ix, jb, B.seq + B.seq_off[jb.mate_read]);
            }
            pool_next += __popc(__ballot_sync(0xffffffffu, take));
        }
        if (!__ballot_sync(0xffffffffu, L.state != Machine::IDLE)) break;      // nothing in flight and nothing left to draw
        if (L.state == Machine::PASS_END && L.end_pass()) {
            SwResult r = L.out;
            if (L.err) r.score = -0x7fffffff;                                   // not this kernel's case: the pair takes the warp path
            res[mine] = r;
        }
        __syncwarp();
        if (L.state == Machine::INIT) L.init_step(64);
        __syncwarp();
        if (L.state == Machine::ROWS) L.step(opt, chunk);
        __syncwarp();
    }
}

// SMEM_LISTS: the pairs that reach this kernel are the ones with long region lists (a read of a two-letter alphabet under the
// wrong conversion pattern has a hundred regions and as many rescue calls), and mem_matesw re-sorts the mate's list after
// every call (mem_sort_dedup_patch, two unstable sorts whose permutations are result-visible): a serial chain of 88-byte
// swaps. With the two lists of the pair in shared memory that chain runs at shared-memory latency: one pair per block of
// one warp, lane 0 alone working (the lists of a pair take tens of KB: three pairs per SM either way).
template <bool SMEM_LISTS>
__global__ void __launch_bounds__(32, 16) k_rescue_replay(Opt opt, IndexView ix, BatchDev B, FinalLayout L, uint8_t *scratch, const int32_t *heavy, int n_heavy,
                                                          const RescueJob *jobs, const SwResult *res, const uint32_t *job_off, int32_t *heavy2, int *n_heavy2, int *ctr,
                                                          unsigned long long *slowest)
{
    extern __shared__ __align__(16) uint8_t rr_lists[];
    const unsigned lane = threadIdx.x & 31;
    const unsigned w = SMEM_LISTS ? blockIdx.x : blockIdx.x * blockDim.x + threadIdx.x, nw = SMEM_LISTS ? gridDim.x : gridDim.x * blockDim.x;   // scratch block of this worker
    FinalWS ws; AlnReg *wregs;
    make_ws(L, scratch + (size_t)w * L.total, ws, wregs);
    if (SMEM_LISTS) {
        AlnReg *lists = reinterpret_cast<AlnReg *>(rr_lists);
        for (;;) {
            unsigned k = 0;
            if (lane == 0) k = (unsigned)atomicAdd(ctr, 1);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= (unsigned)n_heavy) return;
            const RescuePre pre = {jobs + job_off[k], res + job_off[k], (int)(job_off[k + 1] - job_off[k])};
            bool lane_ok = true;
            for (int q = lane; q < pre.n; q += 32) lane_ok = lane_ok && pre.res[q].score != -0x7fffffff;
            if (!__all_sync(0xffffffffu, lane_ok)) { if (lane == 0) heavy2[atomicAdd(n_heavy2, 1)] = heavy[k]; continue; }
            const long long t0 = slowest ? clock64() : 0;
            stage_final_pe_replay_warp(opt, ix, B, heavy[k], ws, lists, pre);
            if (slowest && lane == 0) {
                const unsigned long long dt = (unsigned long long)(clock64() - t0);
                atomicMax(slowest, (dt >> 10) << 32 | (unsigned long long)k);
                atomicAdd(slowest + 1, dt >> 10);
            }
        }
    }
    for (unsigned base = warp_next_group(ctr, (w & ~31u) - nw, nw); base < (unsigned)n_heavy; base = warp_next_group(ctr, base, nw)) {
        const unsigned k = base + lane;
        if (k >= (unsigned)n_heavy) continue;
        const RescuePre pre = {jobs + job_off[k], res + job_off[k], (int)(job_off[k + 1] - job_off[k])};
        bool lane_ok = true;
        for (int q = 0; q < pre.n; ++q) lane_ok = lane_ok && pre.res[q].score != -0x7fffffff;
        if (!lane_ok) { heavy2[atomicAdd(n_heavy2, 1)] = heavy[k]; continue; }
        const long long t0 = slowest ? clock64() : 0;
        stage_final_pe(opt, ix, B, heavy[k], ws, wregs, false, nullptr, &pre);
        if (slowest) {   // diagnostics (BSB_DEBUG_STATS): the slowest pair of the launch, and the cycles of all pairs together
            const unsigned long long dt = (unsigned long long)(clock64() - t0);
            atomicMax(slowest, (dt >> 10) << 32 | (unsigned long long)k);
            atomicAdd(slowest + 1, dt >> 10);
        }
    }
}

// SAM text on the device (bsb_sam.h): sizes, then (after a scan) the bytes, one thread per entry
__global__ void __launch_bounds__(128) k_sam_count(SamView v, int n, int is_pe, uint32_t *len, SamStats *stats)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SamCount o;
    SamStats st;
    sam_entry(o, v, i, is_pe != 0, st);
    len[i] = (uint32_t)o.n;
    stats[i] = st;
}

__global__ void __launch_bounds__(128) k_sam_write(SamView v, int n, int is_pe, const uint32_t *off, char *text)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SamWriteDev o = {text + off[i], 0, 0};
    SamStats st;
    sam_entry(o, v, i, is_pe != 0, st);
    o.finish();
}

// ------------------------------------------------------------------------------------------------
// BAM on the device: the arbiter, the records (bsb_bam.h), the BGZF blocks (bsb_deflate.h)
// ------------------------------------------------------------------------------------------------
// counters of a batch: [0, 8) MapCounters, 8 = first defect ((entry << 8 | code) + 1), 9 = largest entry in bytes, 10 = records
__global__ void __launch_bounds__(128) k_bam_arbiter(ArbiterView a, uint8_t *code, unsigned long long *ctr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    MapCounters c;
    memset(&c, 0, sizeof c);
    if (i < a.n && a.is_head(i)) bam_arbitrate(a, i, code, c);     // one thread per read-name group (its first entry)
    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(&c);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, (unsigned)f[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ctr + k, (unsigned long long)v);
    }
}

__global__ void __launch_bounds__(128) k_bam_count(BamView v, uint32_t *len, unsigned long long *ctr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bytes = 0, nr = 0;
    if (i < v.a.n) {
        SamCount c;
        int r = 0;
        const int rc = bam_entry(c, v, i, false, &r);
        bytes = (unsigned)c.n; nr = (unsigned)r;
        len[i] = bytes;
        if (rc) atomicCAS(ctr + 8, 0ull, ((unsigned long long)i << 8 | (unsigned)rc) + 1);
    }
    const unsigned mx = __reduce_max_sync(0xffffffffu, bytes), sum = __reduce_add_sync(0xffffffffu, nr);
    if ((threadIdx.x & 31) == 0) { atomicMax(ctr + 9, (unsigned long long)mx); if (sum) atomicAdd(ctr + 10, (unsigned long long)sum); }
}

__global__ void __launch_bounds__(128) k_bam_write(BamView v, const uint32_t *off, char *raw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.a.n || off[i + 1] == off[i]) return;
    SamWriteDev o = {raw + off[i], 0, 0};
    bam_entry(o, v, i, true);                                       // (defects were reported by the sizing pass)
    o.finish();
}

__global__ void k_bam_cuts(const uint32_t *off, int n, uint32_t quantum, int fixed, int nblk, uint32_t *cut)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nblk) return;
    const uint64_t target = (uint64_t)k * quantum;
    cut[k] = fixed ? (uint32_t)(target < off[n] ? target : off[n]) : bam_block_cut(off, n, target);
}

struct XDev {   // bsb_deflate.h's execution policy on the device: one phase = a strided loop over the block + a barrier
    unsigned long long *prof; long long last;      // BSB_DF_PROFILE: cycles per phase, summed over the thread blocks (thread 0's clock)
    __host__ __device__ __forceinline__ void tick(int k)
    {
#if defined(__CUDA_ARCH__)
        if (prof && threadIdx.x == 0) { const long long now = clock64(); atomicAdd(prof + k, (unsigned long long)(now - last)); last = now; }
#endif
    }
    template <class F> __host__ __device__ __forceinline__ void par(int n, F f)
    {
#if defined(__CUDA_ARCH__)
        for (int i = threadIdx.x; i < n; i += blockDim.x) f(i);
        __syncthreads();
#endif
    }
    template <class F> __host__ __device__ __forceinline__ void par1(F f)     // one item per thread (blockDim.x == DF_CH)
    {
#if defined(__CUDA_ARCH__)
        f((int)threadIdx.x);
        __syncthreads();
#endif
    }
    template <class F> __host__ __device__ __forceinline__ void wpar1(F f)    // ... with a warp barrier only
    {
#if defined(__CUDA_ARCH__)
        f((int)threadIdx.x);
        __syncwarp();
#endif
    }
    __host__ __device__ __forceinline__ void sync()
    {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
    __host__ __device__ __forceinline__ void atomic_or(uint32_t *p, uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicOr(p, v);
#endif
    }
    __host__ __device__ __forceinline__ void atomic_xor(uint32_t *p, uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicXor(p, v);
#endif
    }
    __host__ __device__ __forceinline__ void atomic_add(uint32_t *p, uint32_t v)
    {
#if defined(__CUDA_ARCH__)
        atomicAdd(p, v);
#endif
    }
    __host__ __device__ __forceinline__ void atomic_max16(uint32_t *w, uint32_t idx, uint32_t v)   // 16-bit entry idx of a word array
    {
#if defined(__CUDA_ARCH__)
        uint32_t *p = w + (idx >> 1);
        const int sh = (int)(idx & 1) << 4;
        uint32_t old = *p;
        while (v > (old >> sh & 0xffffu)) {
            const uint32_t seen = atomicCAS(p, old, (old & ~(0xffffu << sh)) | v << sh);
            if (seen == old) break;
            old = seen;
        }
#endif
    }
};

// One thread block per BGZF block, persistent over the batch's blocks. tok: BGZF_MAX_IN + 8 words of scratch per thread block.
__global__ void __launch_bounds__(DF_CH, 2) k_bgzf_deflate(const uint8_t *raw, const uint32_t *cut, int nblk, uint8_t *slots, uint32_t *len, uint32_t *tok, unsigned long long *prof)
{
    extern __shared__ __align__(16) unsigned char df_smem[];       // sizeof(DeflateShared): the input block, the hash table, the code tables
    DeflateShared &S = *reinterpret_cast<DeflateShared *>(df_smem);
    XDev x;
    x.prof = prof; x.last = clock64();
    uint32_t *my_tok = tok + (size_t)blockIdx.x * (BGZF_MAX_IN + 8);
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int n = (int)(cut[blk + 1] - cut[blk]);
        uint32_t total = 0;
        if (n > 0 && n <= BGZF_MAX_IN) total = bgzf_block(x, S, raw + cut[blk], n, slots + (size_t)blk * BGZF_SLOT, my_tok);
        if (threadIdx.x == 0) len[blk] = n > BGZF_MAX_IN ? 0xffffffffu : total;   // (a cut larger than a block cannot happen; it would be reported)
    }
}

// the finished blocks, back to back
__global__ void __launch_bounds__(256) k_bgzf_gather(const uint8_t *slots, const uint32_t *len, const uint32_t *off, int nblk, uint8_t *dense)
{
    const int blk = blockIdx.x;
    const uint8_t *src = slots + (size_t)blk * BGZF_SLOT;
    uint8_t *dst = dense + off[blk];
    const uint32_t n = len[blk];
    uint32_t i = threadIdx.x;
    const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);   // bytes up to the first 16-byte aligned destination
    if (n >= 64) {
        if (i < head) dst[i] = src[i];
        const uint32_t n16 = (n - head) >> 4;
        for (uint32_t k = i; k < n16; k += blockDim.x) {
            const uint8_t *s = src + head + (k << 4);
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = df_load32(s + 4 * q);
            *reinterpret_cast<uint4 *>(dst + head + (k << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (uint32_t k = head + (n16 << 4) + i; k < n; k += blockDim.x) dst[k] = src[k];
    } else for (; i < n; i += blockDim.x) dst[i] = src[i];
}

// Index-load time: expands the reference's SA sample (every sa_intv-th rank) into the full suffix array in
// HBM. SA values do not depend on the sampling rate, so lookups become one 4-byte load
// instead of ~31 dependent 64-byte LF steps. One thread per sample walks LF until the next sampled rank.
// sa_hi (optional): bits 32..39 of every entry, for texts of 2^32 symbols or more.
__global__ void k_dense_sa(IndexView ix, uint64_t n_sa, uint32_t *sa32, uint8_t *sa_hi)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sa) return;
    uint64_t k = j * (uint64_t)ix.sa_intv;
    uint64_t v = j == 0 ? ix.seq_len : ix.sa[j];
    sa32[k] = j == 0 ? 0xffffffffu : (uint32_t)v;
    if (sa_hi) sa_hi[k] = j == 0 ? 0xff : (uint8_t)(v >> 32);
    const uint64_t mask = (uint64_t)ix.sa_intv - 1;
    for (;;) {
        k = fm_lf(ix, k);
        --v;
        if ((k & mask) == 0) break;
        sa32[k] = (uint32_t)v;
        if (sa_hi) sa_hi[k] = (uint8_t)(v >> 32);
    }
}

__global__ void k_clear_code(int32_t *a, int n, int32_t code)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] == code) a[i] = 0;
}

__global__ void k_max_i32(const int32_t *a, int n, int32_t *out)
{
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, a[i]);
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------------
// page-locked host buffers for everything that crosses PCIe (installed once, at library load)
// ------------------------------------------------------------------------------------------------
// Page-locking is expensive (~0.3 ms/MB), so released blocks go to a process-wide pool and are handed out
// again: the second run of a process (and every recycled batch) reuses the blocks of the first.
struct PinnedPool {
    std::mutex m;
    std::vector<std::pair<uint8_t *, size_t>> free_blocks;   // (block incl. 64-byte header, payload capacity)
    std::map<size_t, int> created;                            // page-locked blocks that exist per size class, handed out or not
};
static PinnedPool &pinned_pool() { static PinnedPool *p = new PinnedPool; return *p; }

static size_t pinned_class(size_t n)
{   // coarse size classes so that buffers of "about the same" size are interchangeable
    if (n > (16u << 20)) n = (n + (32u << 20) - 1) / (32u << 20) * (32u << 20);
    else if (n > (1u << 20)) n = (n + (4u << 20) - 1) / (4u << 20) * (4u << 20);
    return n;
}

static void *pinned_alloc(size_t n)
{
    n = pinned_class(n);
    {
        PinnedPool &P = pinned_pool();
        std::lock_guard<std::mutex> l(P.m);
        int best = -1;
        for (int i = 0; i < (int)P.free_blocks.size(); ++i)
            if (P.free_blocks[i].second >= n && P.free_blocks[i].second <= 2 * n + (1 << 20) &&
                (best < 0 || P.free_blocks[i].second < P.free_blocks[best].second)) best = i;
        if (best >= 0) {
            uint8_t *p = P.free_blocks[best].first;
            P.free_blocks.erase(P.free_blocks.begin() + best);
            return p + 64;
        }
    }
    // 64-byte header: how the block was obtained + its capacity
    void *p = nullptr;
    static const bool dbg_pin = getenv("BSB_DEBUG_PINNED") != nullptr;
    const auto t_pin = std::chrono::steady_clock::now();
    const cudaError_t pin_rc = cudaHostAlloc(&p, n + 64, cudaHostAllocPortable);
    if (dbg_pin) fprintf(stderr, "[D::pinned] cudaHostAlloc %zu MB on demand: %.1f ms\n", n >> 20, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pin).count());
    if (pin_rc == cudaSuccess && p) {
        ((uint64_t *)p)[0] = 0x50494e4e45445f5full; ((uint64_t *)p)[1] = n;
        PinnedPool &P = pinned_pool();
        std::lock_guard<std::mutex> l(P.m);
        ++P.created[n];
        return (uint8_t *)p + 64;
    }
    cudaGetLastError();
    p = malloc(n + 64);
    if (!p) throw std::bad_alloc();
    ((uint64_t *)p)[0] = 0; ((uint64_t *)p)[1] = n;
    return (uint8_t *)p + 64;
}
static void pinned_release(void *q)
{
    uint8_t *p = (uint8_t *)q - 64;
    if (((uint64_t *)p)[0] == 0x50494e4e45445f5full) {
        PinnedPool &P = pinned_pool();
        std::lock_guard<std::mutex> l(P.m);
        P.free_blocks.emplace_back(p, (size_t)((uint64_t *)p)[1]);
    } else free(p);
}
// Page-locks `count` blocks of the class of `bytes` ahead of need (called off the critical path once the first batch
// has shown what sizes a run uses), so that later batches never wait ~30 ms for cudaHostAlloc.
static void pinned_prefill_class(size_t n, int count, bool free_only);
static void pinned_prefill(const size_t *bytes, const int *count, int n_req, bool free_only)
{
    std::map<size_t, int> want;
    for (int i = 0; i < n_req; ++i) want[pinned_class(bytes[i])] += count[i];
    for (const auto &kv : want) pinned_prefill_class(kv.first, kv.second, free_only);
}
static void pinned_prefill_class(size_t n, int count, bool free_only)
{
    PinnedPool &P = pinned_pool();
    // blocks of this class that exist already, whether in the pool or handed out to the batches in flight (a later run of
    // the same process finds the blocks of the first: nothing is page-locked again)
    int have = 0;
    {
        std::lock_guard<std::mutex> l(P.m);
        if (free_only) { for (auto &b : P.free_blocks) if (b.second == n) ++have; }
        else have = P.created[n];
    }
    for (; have < count; ++have) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, n + 64, cudaHostAllocPortable) != cudaSuccess || !p) { cudaGetLastError(); return; }   // portable: every device of a multi-GPU run copies from it
        ((uint64_t *)p)[0] = 0x50494e4e45445f5full; ((uint64_t *)p)[1] = n;
        std::lock_guard<std::mutex> l(P.m);
        P.free_blocks.emplace_back((uint8_t *)p, n);
        ++P.created[n];
    }
}
struct InstallPinnedHooks { InstallPinnedHooks() { g_host_alloc.alloc = pinned_alloc; g_host_alloc.release = pinned_release; g_host_alloc.prefill = pinned_prefill; } };
static InstallPinnedHooks g_install_pinned_hooks;

// ------------------------------------------------------------------------------------------------
// CudaAligner
// ------------------------------------------------------------------------------------------------
// per-batch device state: one stream and its buffers. A CudaAligner owns several so that several batches can be in
// flight on the GPU at once (each driven by its own host thread): the kernels of one batch fill the tails and the
// host round trips of the other.
struct BatchCtx {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[13];   // 0..9 stage boundaries, 10 = selection kernel done, 11 = SAM text written, 12 = BGZF blocks gathered
    cudaEvent_t ev_wait = nullptr;   // blocking-sync event: the host thread sleeps instead of spinning on a core
    DevBuf<char> d_bases; DevBuf<uint32_t> d_seq_off; DevBuf<uint8_t> d_pattern, d_seq, d_oseq;
    DevBuf<Intv> d_intv;
    DevBuf<uint64_t> d_seq4; DevBuf<uint4> d_spill; DevBuf<int32_t> d_cnt_ab;
    DevBuf<int32_t> d_n_intv, d_l_rep, d_n_seed, d_n_chain, d_n_regs, d_err, d_misc;
    DevBuf<uint32_t> d_seed_off;
    DevBuf<uint8_t> d_cub;
    DevBuf<Seed> d_seeds, d_cseeds; DevBuf<int32_t> d_next, d_tmp, d_chain_aux; DevBuf<Chain> d_pool, d_chains;
    DevBuf<uint64_t> d_srt; DevBuf<AlnReg> d_regs; DevBuf<BtNode> d_nodes;
    DevBuf<int32_t> d_eh;
    DevBuf<int8_t> d_pe_dir; DevBuf<int64_t> d_pe_isize;
    DevBuf<double> d_pair;
    DevBuf<ReadOut> d_out; DevBuf<uint8_t> d_arena, d_final_scratch, d_zbuf;
    DevBuf<AlnTask> d_tasks; DevBuf<unsigned int> d_ntasks;
    DevBuf<uint32_t> d_task_cigar; DevBuf<int32_t> d_task_ncig; DevBuf<char> d_task_text;
    DevBuf<int32_t> d_heavy, d_heavy2; DevBuf<uint32_t> d_job_cnt, d_job_off; DevBuf<RescueJob> d_jobs; DevBuf<SwResult> d_job_res; DevBuf<uint64_t> d_blist;
    // device-side SAM text
    DevBuf<char> d_names, d_qual, d_text, d_rg; DevBuf<uint32_t> d_name_off, d_text_len, d_text_off; DevBuf<uint8_t> d_has_qual; DevBuf<SamStats> d_stats;
    // device-side BAM: arbiter decisions, record bytes, BGZF blocks
    DevBuf<uint8_t> d_first, d_rgrp, d_bam_code, d_bam_raw, d_bgzf_slots, d_bgzf_dense;
    DevBuf<uint32_t> d_bam_len, d_bam_off, d_bam_cut, d_bgzf_len, d_bgzf_off, d_bam_tok;
    DevBuf<unsigned long long> d_bam_ctr;
    size_t task_cap = 0;
    DevBuf<unsigned long long> d_used, d_work;   // d_work: counted work of the batch (FM extensions, two-block extensions, extension DP cells)
    size_t arena_cap = 0;
    int intv_cap_hint = 0;
    int fin_scale = 1;          // growth of the finalisation scratch (see make_layout in align())
    int last_heavy = -1, last_pairs = 0;   // pairs queued for rescue by this context's previous batch, of how many: picks the rescue path without a host round trip
    long launches = 0;
    bool ready = false;
    void init()
    {
        if (ready) return;
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto &e : ev) CK(cudaEventCreate(&e));
        CK(cudaEventCreateWithFlags(&ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
        d_used.ensure(1); d_misc.ensure(32); d_ntasks.ensure(1); d_work.ensure(8);
        ready = true;
    }
    ~BatchCtx()
    {
        if (!ready) return;
        for (auto &e : ev) cudaEventDestroy(e);
        if (ev_wait) cudaEventDestroy(ev_wait);
        if (st) cudaStreamDestroy(st);
    }
    void wait()
    {
        static const bool spin = getenv("BSB_SPIN") != nullptr;
        if (spin) { CK(cudaStreamSynchronize(st)); return; }
        CK(cudaEventRecord(ev_wait, st));
        CK(cudaEventSynchronize(ev_wait));
    }
};

struct CudaAligner::Impl {
    int device = 0, n_sm = 0;
    // resident index
    DevBuf<uint32_t> d_bwt, d_sa32, d_occ32; DevBuf<uint64_t> d_sa; DevBuf<uint8_t> d_pac, d_opac, d_sa_hi; DevBuf<Ann> d_anns;
    IndexView ix;
    DevBuf<double> d_log;
    std::vector<double> log_tab;
    DevBuf<char> d_ctg_text; DevBuf<uint32_t> d_ctg_name_off, d_ctg_anno_off; DevBuf<uint8_t> d_ctg_is_crick, d_ctg_sign;   // SAM formatter's contig table
    DevBuf<int32_t> d_ctg_sorted; int n_ctg = 0, df_per_sm = 0;                                                                          // + the BAM encoder's name lookup
    bool any_alt = false;
    long launches = 0;        // index-load kernels
    BatchCtx ctx[CudaAligner::kSlots];
    std::mutex init_m;
};

CudaAligner::CudaAligner(const HostIndex &idx, int device) : im_(new Impl)
{
    Impl &m = *im_;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw std::runtime_error(std::string("[E::bsbolt_b200] no CUDA device available (") + cudaGetErrorString(e) + "); this aligner has no CPU fallback");
    if (device < 0 || device >= ndev) throw std::runtime_error("[E::bsbolt_b200] invalid CUDA device index " + std::to_string(device));
    m.device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    m.n_sm = prop.multiProcessorCount;
    CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384));
    // index -> HBM (reference layout, see bsb_index.h)
    m.d_bwt.ensure(idx.bwt.size()); CK(cudaMemcpy(m.d_bwt.p, idx.bwt.data(), idx.bwt.size() * 4, cudaMemcpyHostToDevice));
    m.d_sa.ensure(idx.sa.size()); CK(cudaMemcpy(m.d_sa.p, idx.sa.data(), idx.sa.size() * 8, cudaMemcpyHostToDevice));
    m.d_pac.ensure(idx.pac.size()); CK(cudaMemcpy(m.d_pac.p, idx.pac.data(), idx.pac.size(), cudaMemcpyHostToDevice));
    m.d_opac.ensure(idx.opac.size()); CK(cudaMemcpy(m.d_opac.p, idx.opac.data(), idx.opac.size(), cudaMemcpyHostToDevice));
    m.d_anns.ensure(idx.anns.size()); CK(cudaMemcpy(m.d_anns.p, idx.anns.data(), idx.anns.size() * sizeof(Ann), cudaMemcpyHostToDevice));
    m.ix = idx.host_view();
    m.ix.bwt = m.d_bwt.p; m.ix.sa = m.d_sa.p; m.ix.pac = m.d_pac.p; m.ix.opac = m.d_opac.p; m.ix.anns = m.d_anns.p;
    m.ix.sa32 = nullptr; m.ix.sa32_intv = 0; m.ix.occ32 = nullptr; m.ix.sa_hi = nullptr;
    if (idx.seq_len + 1 < (1ull << 32) && !getenv("BSB_REF_BLOCKS")) {   // sector-sized occ blocks for seeding
        const uint64_t nb32 = (uint64_t)(idx.bwt.size() / 16) * 2;
        m.d_occ32.ensure(nb32 * 8);
        k_make_occ32<<<(unsigned)((nb32 + 255) / 256), 256>>>(m.d_bwt.p, nb32, m.d_occ32.p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        m.ix.occ32 = m.d_occ32.p;
        ++m.launches;
    }
    // full SA resident in HBM: 4 B/rank below 2^32 symbols, else 5 B/rank (40-bit entries) when it fits beside the rest of
    // the index with room left for the batches (a 12.4 G-symbol human-scale text: 62 GB of the 180 GB)
    const bool wide_sa = idx.seq_len + 1 >= (1ull << 32) || getenv("BSB_DENSE_SA40");
    bool dense_sa = !getenv("BSB_SAMPLED_SA");
    if (dense_sa && wide_sa) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = (size_t)(idx.seq_len + 1) * 5, reserve = (size_t)32 << 30;
        if (idx.seq_len >= (1ull << 40) || (free_b < need + reserve && !getenv("BSB_DENSE_SA40"))) dense_sa = false;
    }
    if (dense_sa) {
        m.d_sa32.ensure(idx.seq_len + 1);
        if (wide_sa) m.d_sa_hi.ensure(idx.seq_len + 1);
        k_dense_sa<<<(unsigned)((idx.n_sa + 127) / 128), 128>>>(m.ix, idx.n_sa, m.d_sa32.p, wide_sa ? m.d_sa_hi.p : nullptr);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        m.ix.sa32 = m.d_sa32.p; m.ix.sa32_intv = 1; m.ix.sa_hi = wide_sa ? m.d_sa_hi.p : nullptr;
        ++m.launches;
    }
    {   // contig table of the SAM formatter
        auto up = [](auto &dst, const auto &src) { dst.ensure(src.size() + 1); if (!src.empty()) CK(cudaMemcpy(dst.p, src.data(), src.size() * sizeof(src[0]), cudaMemcpyHostToDevice)); };
        up(m.d_ctg_text, idx.ctg_text); up(m.d_ctg_name_off, idx.ctg_name_off); up(m.d_ctg_anno_off, idx.ctg_anno_off);
        up(m.d_ctg_is_crick, idx.ctg_is_crick); up(m.d_ctg_sign, idx.ctg_sign); up(m.d_ctg_sorted, idx.ctg_sorted);
        m.n_ctg = (int)idx.ctg_sorted.size();
        m.any_alt = idx.any_alt;
    }
    build_log_table(m.log_tab, 65536);
    m.d_log.ensure(m.log_tab.size());
    CK(cudaMemcpy(m.d_log.p, m.log_tab.data(), m.log_tab.size() * 8, cudaMemcpyHostToDevice));
}

CudaAligner::~CudaAligner()
{
    if (!im_) return;
    cudaSetDevice(im_->device);
    delete im_;
}

struct DeviceInput { DevBuf<char> bases; DevBuf<uint32_t> seq_off; DevBuf<uint8_t> pattern; };

void CudaAligner::preload(ReadBatch &b, int)
{
    CK(cudaSetDevice(im_->device));
    DeviceInput *d = new DeviceInput;
    const size_t nb = b.bases.size();
    d->bases.ensure(nb + 16); d->seq_off.ensure(b.n + 1); d->pattern.ensure(b.n + 1);
    CK(cudaMemcpy(d->bases.p, b.bases.data(), nb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->seq_off.p, b.seq_off.data(), (size_t)(b.n + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->pattern.p, b.pattern.data(), (size_t)b.n, cudaMemcpyHostToDevice));
    b.dev_input = d;
}

void CudaAligner::unload(ReadBatch &b)
{
    if (!b.dev_input) return;
    cudaSetDevice(im_->device);
    delete static_cast<DeviceInput *>(b.dev_input);
    b.dev_input = nullptr;
}

long CudaAligner::kernel_launches() const
{
    long n = im_->launches;
    for (const BatchCtx &c : im_->ctx) n += c.launches;
    return n;
}
int CudaAligner::device() const { return im_->device; }
size_t CudaAligner::index_bytes() const
{
    const Impl &m = *im_;
    return m.d_bwt.cap * 4 + m.d_occ32.cap * 4 + m.d_sa.cap * 8 + m.d_sa32.cap * 4 + m.d_sa_hi.cap + m.d_pac.cap + m.d_opac.cap;
}

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static const int kMaxReadLen = 1200;

// cudaFuncAttributeMaxDynamicSharedMemorySize is one value per kernel and device, shared by the device slots that launch
// concurrently with different batch shapes: it is only ever raised, under a lock, so that a slot with a smaller batch cannot
// lower it between another slot's request and launch.
template <class K>
static void raise_dynamic_smem(K kernel, int bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, int> cur;
    if (bytes <= 48 * 1024) return;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> l(mu);
    int &c = cur[std::make_pair(dev, (const void *)kernel)];
    if (bytes > c) { CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); c = bytes; }
}

void CudaAligner::align(const Opt &opt_in, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int slot)
{
    Impl &I = *im_;
    if (slot < 0 || slot >= kSlots) throw std::runtime_error("[E::bsbolt_b200] invalid batch slot");
    CK(cudaSetDevice(I.device));
    BatchCtx &m = I.ctx[slot];
    { std::lock_guard<std::mutex> l(I.init_m); m.init(); }
    cudaStream_t st = m.st;
    const Opt opt = opt_in;
    const int n = b.n;
    const bool pe = (opt.flag & F_PE) != 0;
    if (pe && (n & 1)) throw std::runtime_error("[E::bsbolt_b200] paired-end batch with an odd number of reads");
    const size_t nb = b.bases.size();
    int max_len = 0;
    for (int i = 0; i < n; ++i) max_len = std::max(max_len, b.len(i));
    // the warp kernels keep a DP row, the query and the CIGAR/MD scratch of a read in shared memory (36 bytes x length per block)
    // offsets into the batch's bases and into its SAM text (3 to 4 bytes of text per base) are 32-bit on the device
    if (nb > 900000000ull) throw std::runtime_error("[E::bsbolt_b200] a batch of " + std::to_string(nb) + " bases: batches above 900 Mbp are not supported (lower -K, or -t when -K is not given)");
    if (max_len > kMaxReadLen) throw std::runtime_error("[E::bsbolt_b200] reads longer than " + std::to_string(kMaxReadLen) + " bp do not fit the per-warp shared-memory tiles of this build");
    for (int k = 0; k < 8; ++k) out.ms_stage[k] = 0;
    out.n_rescue_jobs = out.n_rescue_pairs = 0;
    const bool dbg = getenv("BSB_DEBUG_TIMELINE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    static const auto t_epoch = std::chrono::steady_clock::now();
    auto T = [&](const char *tag) {
        if (dbg) fprintf(stderr, "[D::timeline] slot %d %-14s %8.2f ms (abs %9.2f)\n", slot, tag, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_epoch).count());
    };

    // ---- H2D ----
    CK(cudaEventRecord(m.ev[0], st));
    m.d_bases.ensure(nb + 16); m.d_seq.ensure(nb + 16); m.d_oseq.ensure(nb + 16);
    m.d_seq_off.ensure(n + 1); m.d_pattern.ensure(n + 1);
    const DeviceInput *pre = static_cast<const DeviceInput *>(b.dev_input);
    if (!pre) {
        CK(cudaMemcpyAsync(m.d_bases.p, b.bases.data(), nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m.d_seq_off.p, b.seq_off.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m.d_pattern.p, b.pattern.data(), (size_t)n, cudaMemcpyHostToDevice, st));
    }
    const bool dev_text = out.want_text && !I.any_alt && b.comments.empty() && !getenv("BSB_HOST_FORMAT");
    if (dev_text) {   // what the device formatter reads besides the results: names, qualities
        m.d_names.ensure(b.names.size() + 16); m.d_name_off.ensure(n + 2); m.d_qual.ensure(nb + 16); m.d_has_qual.ensure(n + 1);
        CK(cudaMemcpyAsync(m.d_names.p, b.names.data(), b.names.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m.d_name_off.p, b.name_off.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m.d_qual.p, b.qual.data(), nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m.d_has_qual.p, b.has_qual.data(), (size_t)n, cudaMemcpyHostToDevice, st));
    }
    CK(cudaEventRecord(m.ev[1], st));
    T("h2d_enq");

    BatchDev B;
    memset(&B, 0, sizeof B);
    B.n = n; B.is_pe = pe; B.n_processed = n_processed;
    B.bases = m.d_bases.p; B.seq_off = m.d_seq_off.p; B.pattern = m.d_pattern.p; B.seq = m.d_seq.p; B.oseq = m.d_oseq.p;
    if (pre) { B.bases = pre->bases.p; B.seq_off = pre->seq_off.p; B.pattern = pre->pattern.p; }
    m.d_n_intv.ensure(n + 1); m.d_l_rep.ensure(n + 1); m.d_n_seed.ensure(n + 1); m.d_n_chain.ensure(n + 1); m.d_n_regs.ensure(n + 1);
    m.d_err.ensure(n + 1); m.d_seed_off.ensure(n + 2);
    B.n_intv = m.d_n_intv.p; B.l_rep = m.d_l_rep.p; B.n_seed = m.d_n_seed.p; B.n_chain = m.d_n_chain.p; B.n_regs = m.d_n_regs.p;
    B.err = m.d_err.p;

    CK(cudaMemsetAsync(m.d_work.p, 0, 4 * sizeof(unsigned long long), st));
    // ---- K1 ----
    if (nb) { k_convert<<<cdiv(cdiv(nb, 16), 256), 256, 0, st>>>(B, (uint32_t)nb); ++m.launches; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(m.ev[2], st));

    // ---- K2 (retry with a larger interval capacity on overflow) ----
    B.intv_cap = std::max(std::max(256, 2 * max_len), m.intv_cap_hint);
    auto env_int = [](const char *k, int d) { const char *v = getenv(k); return v ? atoi(v) : d; };
    const int s3_scap = I.ix.occ32 ? env_int("BSB_S3_SCAP", 16) : 16, s3_total = max_len + 1, s3_bps = env_int("BSB_S3_BPS", I.ix.occ32 ? 12 : 10), s3_blocks = I.n_sm * s3_bps;
    const uint32_t n_words = (uint32_t)(nb >> 4) + (uint32_t)n + 1;
    if (n) {
        m.d_seq4.ensure(n_words + 1);
        m.d_spill.ensure((size_t)s3_blocks * SEED3_BLOCK * (size_t)s3_total + 1);
        m.d_cnt_ab.ensure(2 * (size_t)n + 2);
        k_pack4<<<cdiv(n_words, 256), 256, 0, st>>>(B, m.d_seq4.p, n_words); ++m.launches;
        CK(cudaGetLastError());
    }
    for (;;) {
        m.d_intv.ensure((size_t)n * B.intv_cap);
        B.intv = m.d_intv.p;
        CK(cudaMemsetAsync(m.d_err.p, 0, (size_t)(n + 1) * 4, st));
        CK(cudaMemsetAsync(m.d_n_seed.p, 0, (size_t)(n + 1) * 4, st));
        CK(cudaMemsetAsync(m.d_misc.p, 0, 16 * 4, st));
        if (n) {
            CK(cudaMemsetAsync(m.d_cnt_ab.p, 0, 2 * (size_t)n * 4, st));
            const size_t s3_smem = (size_t)SEED3_BLOCK * s3_scap * 12;
#define BSB_S3_LAUNCH(SC, MB, CP) k_seed3<SC, MB, CP><<<s3_blocks, SEED3_BLOCK, s3_smem, st>>>(opt, I.ix, B, m.d_seq4.p, m.d_spill.p, s3_total, m.d_misc.p + 8, m.d_cnt_ab.p, m.d_cnt_ab.p + n, m.d_work.p)
            if (!I.ix.occ32) BSB_S3_LAUNCH(16, 10, false);       // >= 2^32-symbol index: reference block layout
            else if (s3_bps > 12) { if (s3_scap == 8) BSB_S3_LAUNCH(8, 16, true); else BSB_S3_LAUNCH(16, 16, true); }
            else if (s3_bps > 10) { if (s3_scap == 8) BSB_S3_LAUNCH(8, 12, true); else if (s3_scap == 32) BSB_S3_LAUNCH(32, 12, true); else BSB_S3_LAUNCH(16, 12, true); }
            else { if (s3_scap == 8) BSB_S3_LAUNCH(8, 10, true); else if (s3_scap == 32) BSB_S3_LAUNCH(32, 10, true); else BSB_S3_LAUNCH(16, 10, true); }
#undef BSB_S3_LAUNCH
            k_seed3_finish<<<cdiv((size_t)n * 32, 128), 128, 0, st>>>(opt, B, m.d_cnt_ab.p, m.d_cnt_ab.p + n);
            m.launches += 2;
        }
        CK(cudaGetLastError());
        k_max_i32<<<I.n_sm, 256, 0, st>>>(m.d_err.p, n, m.d_misc.p); ++m.launches;
        int32_t max_err = 0;
        CK(cudaMemcpyAsync(&max_err, m.d_misc.p, 4, cudaMemcpyDeviceToHost, st));
        m.wait();
        if (max_err == 0) break;
        if (max_err != ERR_INTV_OVERFLOW) throw std::runtime_error("[E::bsbolt_b200] seeding failed with error code " + std::to_string(max_err));
        B.intv_cap *= 2;
        m.intv_cap_hint = B.intv_cap;
        if (B.intv_cap > (1 << 20)) throw std::runtime_error("[E::bsbolt_b200] interval list overflow");
    }
    CK(cudaEventRecord(m.ev[3], st));
    T("seed_done");

    // ---- seed offsets ----
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int32_t *)m.d_n_seed.p, (uint32_t *)m.d_seed_off.p, n + 1, st);
    m.d_cub.ensure(cub_bytes + 16);
    cub::DeviceScan::ExclusiveSum(m.d_cub.p, cub_bytes, (const int32_t *)m.d_n_seed.p, (uint32_t *)m.d_seed_off.p, n + 1, st); ++m.launches;
    uint32_t S = 0;
    CK(cudaMemcpyAsync(&S, m.d_seed_off.p + n, 4, cudaMemcpyDeviceToHost, st));
    m.wait();
    B.seed_off = m.d_seed_off.p;
    T("scan_done");
    if (getenv("BSB_DEBUG_STATS")) { // distribution of per-read seed counts (load-balance diagnostics)
        std::vector<int32_t> ns(n);
        CK(cudaMemcpy(ns.data(), m.d_n_seed.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        std::sort(ns.begin(), ns.end());
        long big = 0; for (int v : ns) if (v > 1000) ++big;
        fprintf(stderr, "[D::seeds] reads %d total %u mean %.1f p50 %d p90 %d p99 %d p99.9 %d max %d; >1000 seeds: %ld reads\n", n, S,
                (double)S / std::max(n, 1), ns[n / 2], ns[(size_t)n * 9 / 10], ns[(size_t)n * 99 / 100], ns[(size_t)n * 999 / 1000], ns[n - 1], big);
    }
    m.d_seeds.ensure(S + 1); m.d_cseeds.ensure(S + 1); m.d_next.ensure(S + 1); m.d_tmp.ensure(S + 1);
    m.d_pool.ensure(S + 1); m.d_chains.ensure(S + 1); m.d_srt.ensure(S + 1); m.d_regs.ensure(S + 1);
    m.d_nodes.ensure(S / 4 + 2 * (size_t)n + 4);
    B.seeds = m.d_seeds.p; B.cseeds = m.d_cseeds.p; B.next = m.d_next.p; B.tmp = m.d_tmp.p;
    B.chain_pool = m.d_pool.p; B.chains = m.d_chains.p; B.srt = m.d_srt.p; B.regs = m.d_regs.p; B.nodes = m.d_nodes.p;

    // ---- K3 ----
    if (S) { k_sa<<<cdiv(S, 128), 128, 0, st>>>(opt, I.ix, B, S); ++m.launches; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(m.ev[4], st));
    // ---- K4 ----
    const bool dyn_sched = getenv("BSB_STATIC_SCHED") == nullptr;
    CK(cudaMemsetAsync(m.d_misc.p + 16, 0, 2 * 4, st));        // read counters of the chaining and the extension kernel
    m.d_chain_aux.ensure(S + 1);
    k_chain_warp<<<I.n_sm * env_int("BSB_CHAIN_BPS", 12), 128, 0, st>>>(opt, I.ix, B, dyn_sched ? m.d_misc.p + 16 : nullptr, m.d_chain_aux.p);
    ++m.launches;
    CK(cudaGetLastError());
    // mem_flt_chained_seeds runs for reads with 5.5 ln(l) <= 0.05 l (l of about 720 and more; the kernel applies the exact
    // per-read test) or, with -W, for reads of 22 W bases and more
    const bool chained_seed_sw = opt.min_chain_weight ? !((double)(1.1f * (float)opt.min_chain_weight) > (double)(0.05f * (float)max_len)) : max_len >= 700;
    if (chained_seed_sw && S) {
        k_seed_sw<<<I.n_sm * 8, 128, 0, st>>>(opt, I.ix, B, I.d_log.p, (int)I.log_tab.size()); ++m.launches;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(m.ev[5], st));
    // ---- K5 ----
    const int max_q = max_len + 8;
    // the lane-per-read machine needs every score in 14 bits, the scoring matrix of bwa_fill_scmat (always the case for
    // matrices built from -A/-B) and a shared-memory row tile per warp; longer reads take the warp-per-read kernel
    bool scmat_std = true;
    for (int i = 0; i < 5 && scmat_std; ++i)
        for (int j = 0; j < 5; ++j) {
            const int want = (i == 4 || j == 4) ? -1 : (i == j ? opt.a : -opt.b);
            if (opt.mat[i * 5 + j] != want) { scmat_std = false; break; }
        }
    // a seed is min_seed_len bases or longer, so no flank (the query of one extension) is longer than the read without it
    const int xl_row_cap = std::max(8, max_len - std::max(opt.min_seed_len, 0) + 3);
    const size_t xl_smem = (size_t)XL_WARPS * 32 * (size_t)xl_row_cap * 4 + 3 * (size_t)(max_q + 2) * 4;
    const long xl_max_score = 2L * max_len * (long)std::max(opt.a, 1) + std::max(opt.pen_clip5, opt.pen_clip3) + 64;
    const bool ext_lanes = !getenv("BSB_EXTEND_WARP") && scmat_std && xl_max_score < (1 << XL_BITS) - 1 && max_q < 4000 && xl_smem <= 100 * 1024 &&
                           opt.a > 0 && opt.b >= 0 && opt.e_ins > 0 && opt.e_del > 0;
    bool use_lanes = ext_lanes;
    int32_t h_misc[2] = {0, 0};
    for (;;) {
    if (use_lanes) {
        raise_dynamic_smem(k_extend_lanes, (int)xl_smem);
        const int xl_bps = std::max(1, std::min(16, (int)((225 * 1024) / (xl_smem + 1024))));
        const int blocks = (int)std::min<size_t>((size_t)cdiv(n, XL_WARPS * 32), (size_t)I.n_sm * xl_bps);
        k_extend_lanes<<<blocks, XL_WARPS * 32, xl_smem, st>>>(opt, I.ix, B, max_q, xl_row_cap, m.d_misc.p + 17, env_int("BSB_XL_CHUNK", 24), env_int("BSB_XL_CTLMASK", 15),
                                                                env_int("BSB_XL_CTLLANES", 10), m.d_work.p);
        const int tail_blocks = (int)std::min<size_t>((size_t)cdiv(n, 128), (size_t)I.n_sm * 8);
        m.d_eh.ensure((size_t)tail_blocks * 128 * 2 * (max_q + 1));
        k_extend_tail<<<tail_blocks, 128, 0, st>>>(opt, I.ix, B, m.d_eh.p, max_q);
        m.launches += 2;
    } else {
        const int wpb = 4;
        const int smem_per_warp = (2 * (max_q + 1) * 4 + max_q + 15) & ~15;
        const int ext_bps = env_int("BSB_EXT_BPS", 5);
        const int32_t *ext_order = nullptr;
        const int blocks = (int)std::min<size_t>((size_t)cdiv(n, wpb), (size_t)I.n_sm * ext_bps);
        m.d_eh.ensure((size_t)blocks * wpb * 32 * 2 * (max_q + 1));   // one (h,e) row per lane for the per-lane tail
        int *ext_ctr = dyn_sched ? m.d_misc.p + 17 : nullptr;
        if (ext_bps > 5) k_extend_warp<8><<<blocks, wpb * 32, wpb * smem_per_warp, st>>>(opt, I.ix, B, m.d_eh.p, max_q, smem_per_warp, ext_order, ext_ctr);
        else k_extend_warp<5><<<blocks, wpb * 32, wpb * smem_per_warp, st>>>(opt, I.ix, B, m.d_eh.p, max_q, smem_per_warp, ext_order, ext_ctr);
        ++m.launches;
    }
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(m.d_misc.p, 0, 16 * 4, st));
    k_max_i32<<<I.n_sm, 256, 0, st>>>(m.d_n_regs.p, n, m.d_misc.p); ++m.launches;
    k_max_i32<<<I.n_sm, 256, 0, st>>>(m.d_err.p, n, m.d_misc.p + 1); ++m.launches;
    CK(cudaMemcpyAsync(h_misc, m.d_misc.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(m.ev[6], st));
    m.wait();
    T("extend_done");
    if (use_lanes && h_misc[1] == ERR_ROW_TILE) {   // a flank longer than the row tile (a seed shorter than -k): warp-per-read form
        k_clear_code<<<cdiv(n, 256), 256, 0, st>>>(m.d_err.p, n, ERR_ROW_TILE); ++m.launches;
        CK(cudaMemsetAsync(m.d_misc.p + 17, 0, 4, st));
        use_lanes = false;
        continue;
    }
    break;
    }
    if (h_misc[1]) throw std::runtime_error("[E::bsbolt_b200] chaining/extension failed with error code " + std::to_string(h_misc[1]));
    const int max_regs = h_misc[0];
    if (getenv("BSB_DEBUG_STATS")) {
        std::vector<int32_t> nr(n), nc(n);
        CK(cudaMemcpy(nr.data(), m.d_n_regs.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(nc.data(), m.d_n_chain.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        std::sort(nr.begin(), nr.end()); std::sort(nc.begin(), nc.end());
        fprintf(stderr, "[D::regs] p50 %d p99 %d p99.9 %d max %d | chains p50 %d p99 %d p99.9 %d max %d\n", nr[n / 2], nr[(size_t)n * 99 / 100],
                nr[(size_t)n * 999 / 1000], nr[n - 1], nc[n / 2], nc[(size_t)n * 99 / 100], nc[(size_t)n * 999 / 1000], nc[n - 1]);
    }

    // ---- insert-size statistics (host, per batch) ----
    std::vector<double> pair_tab;
    if (pe) {
        if (pes0) memcpy(B.pes, pes0, sizeof B.pes);
        else {
            const int np = n >> 1;
            m.d_pe_dir.ensure(np + 1); m.d_pe_isize.ensure(np + 1);
            B.pe_dir = m.d_pe_dir.p; B.pe_isize = m.d_pe_isize.p;
            k_pestat<<<cdiv(np, 128), 128, 0, st>>>(opt, I.ix, B); ++m.launches;
            CK(cudaGetLastError());
            std::vector<int8_t> dir(np); std::vector<int64_t> isz(np);
            CK(cudaMemcpyAsync(dir.data(), m.d_pe_dir.p, np, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(isz.data(), m.d_pe_isize.p, (size_t)np * 8, cudaMemcpyDeviceToHost, st));
            m.wait();
            estimate_pestat(opt, dir, isz, B.pes, verbose, &out.log_text);
        }
        memcpy(out.pes, B.pes, sizeof B.pes);
    }
    build_pair_table(opt, B.pes, pair_tab, B.mt.pair_off);
    m.d_pair.ensure(pair_tab.size());
    CK(cudaMemcpyAsync(m.d_pair.p, pair_tab.data(), pair_tab.size() * 8, cudaMemcpyHostToDevice, st));
    B.mt.pair_tab = m.d_pair.p; B.mt.log_tab = I.d_log.p; B.mt.n_log = (int)I.log_tab.size();
    CK(cudaEventRecord(m.ev[7], st));
    T("pestat_done");

    // ---- K6/K7/K8 ----
    // per-worker scratch of the selection / pairing kernels. The pairing arrays and the sub-optimal list of the rescue
    // Smith-Waterman are sized for ordinary reads; a read that overflows them (ERR_SCRATCH_OVERFLOW: hundreds of candidate
    // pairs in a repeat) makes the stage run again with four times the room -- the reference's kvecs grow the same way
    auto make_layout = [&](int scale) {
        FinalLayout L;
        memset(&L, 0, sizeof L);
        L.max_q = max_q;
        L.reg_cap = max_regs + 4 * opt.max_matesw + 8 + (scale > 16 ? 16 * scale : 0);   // (rescue adds at most 4 regions per call: only the last resort grows this)
        L.pair_cap = 512 * scale; L.sw_cap = max_q + 32; L.sw_b = 1024 * scale; L.wreg_stride = L.reg_cap + opt.max_matesw;
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
        L.eh = take((size_t)2 * (max_q + 1) * 4);
        L.cnt = take((size_t)L.reg_cap * 4); L.has_alt = take(L.reg_cap);
        L.zz = take((size_t)L.reg_cap * 4); L.pv = take((size_t)L.pair_cap * 16); L.pu = take((size_t)L.pair_cap * 16);
        L.sw = take((size_t)4 * L.sw_cap * 4); L.swb = take((size_t)L.sw_b * 8); L.rev = take(max_q);
        L.wregs = take((size_t)2 * L.wreg_stride * sizeof(AlnReg));
        L.total = o;
        return L;
    };
    FinalLayout L = make_layout(m.fin_scale);
    const int fin_block = 32;
    const int items = pe ? n >> 1 : n;
    const int fin_bps = env_int("BSB_FIN_BPS", 16);
    const int fin_workers = (int)std::min<size_t>((size_t)cdiv(std::max(items, 1), fin_block) * fin_block, (size_t)I.n_sm * fin_bps * fin_block);
    const int heavy_blocks = I.n_sm * 4;   // 4 warps each; every warp needs one scratch block like a light worker
    m.d_final_scratch.ensure((size_t)std::max(fin_workers, heavy_blocks * 4) * L.total);
    if (pe) m.d_heavy.ensure((size_t)(n >> 1) + 1);
    m.d_out.ensure(n + 1);
    B.out = m.d_out.p;
    if (m.arena_cap == 0) m.arena_cap = (size_t)n * 640 + (1 << 20);
    if (m.task_cap == 0) m.task_cap = (size_t)n + (size_t)n / 2 + 4096;
    // task kernel geometry
    const int tk_wpb = 4;
    const int tk_blocks = I.n_sm * 8;
    const long z_cap = (long)max_q * (long)(max_q + 2 * (4 * opt.w) + 64);
    m.d_zbuf.ensure((size_t)tk_blocks * tk_wpb * z_cap);
    unsigned long long used = 0;
    unsigned int n_tasks = 0;
    out.reads.resize(n);
    T("final_setup");
    for (;;) {
        m.d_arena.ensure(m.arena_cap);
        m.d_tasks.ensure(m.task_cap);
        unsigned long long init = 8; // offset 0 = "null"
        CK(cudaMemcpyAsync(m.d_used.p, &init, 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(m.d_ntasks.p, 0, 4, st));
        CK(cudaMemsetAsync(m.d_misc.p + 18, 0, 4 * 4, st));   // work counters: selection, rescue, task DP, task finish
        B.arena.base = m.d_arena.p; B.arena.used = m.d_used.p; B.arena.cap = m.arena_cap;
        B.tasks.a = m.d_tasks.p; B.tasks.n = m.d_ntasks.p; B.tasks.cap = (unsigned int)m.task_cap;
        if (items) {
            if (pe) {
                const bool coop = true;    // pairs that need rescue Smith-Waterman are queued for the rescue kernels
                int32_t *heavy = m.d_heavy.p;
                int *n_heavy = m.d_misc.p + 12;
                CK(cudaMemsetAsync(n_heavy, 0, 4, st));
                int *c_fin = dyn_sched ? m.d_misc.p + 18 : nullptr, *c_heavy = dyn_sched ? m.d_misc.p + 19 : nullptr;
                if (fin_bps > 24) k_final_pe<32><<<fin_workers / fin_block, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_fin);
                else if (fin_bps > 16) k_final_pe<24><<<fin_workers / fin_block, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_fin);
                else k_final_pe<16><<<fin_workers / fin_block, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_fin);
                if (coop) {
                    const int hv_smem = 4 * (int)((4 * L.sw_cap + (L.sw_cap + 3) / 4) * 4);
                    raise_dynamic_smem(k_final_pe_heavy, hv_smem);
                    // rescue by jobs: the 8-bit Smith-Waterman (reads below 250 / a bases) with a tile per lane that fits shared memory
                    const int rs_cells = ((max_len + 15) / 16) * 16, rs_words = rs_cells + rs_cells / 8;
                    const size_t rs_smem = (size_t)32 * rs_words * 4;
                    const bool by_jobs = !getenv("BSB_RESCUE_WARP") && (long)max_len * opt.a < 250 && rs_smem <= 96 * 1024;
                    // The job path needs the number of queued pairs on the host. Libraries differ by orders of magnitude in that
                    // number but hardly from batch to batch, so the previous batch of this context decides: few pairs -> the
                    // warp-per-pair kernel straight away (device-side count, no round trip), many -> count, enumerate, run by jobs.
                    const double heavy_expected = m.last_heavy < 0 ? 1e9 : (double)m.last_heavy / std::max(m.last_pairs, 1) * (n >> 1);   // first batch: find out
                    m.last_pairs = n >> 1;
                    if (by_jobs && heavy_expected < env_int("BSB_RESCUE_PAIRS_MIN", 256)) {
                        k_final_pe_heavy<<<heavy_blocks, 128, hv_smem, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_heavy);
                        ++m.launches;
                        CK(cudaMemcpyAsync(&m.last_heavy, n_heavy, 4, cudaMemcpyDeviceToHost, st));   // read at the next wait of this stream
                    } else if (by_jobs) {
                        int h_heavy = 0;
                        CK(cudaMemcpyAsync(&h_heavy, n_heavy, 4, cudaMemcpyDeviceToHost, st));
                        m.wait();
                        m.last_heavy = h_heavy;
                        CK(cudaMemsetAsync(m.d_misc.p + 13, 0, 4, st));            // pairs left to the warp kernel
                        if (h_heavy > 0) {
                            m.d_heavy2.ensure((size_t)h_heavy + 1);
                            m.d_job_cnt.ensure((size_t)h_heavy + 2); m.d_job_off.ensure((size_t)h_heavy + 2);
                            CK(cudaMemsetAsync(m.d_misc.p + 22, 0, 4 * 4, st));    // work counters of the four launches below
                            CK(cudaMemsetAsync(m.d_job_cnt.p + h_heavy, 0, 4, st));
                            const int en_blocks = (int)std::min<size_t>(cdiv(h_heavy, fin_block), (size_t)fin_workers / fin_block);
                            k_rescue_enum<<<en_blocks, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, h_heavy, nullptr, nullptr, m.d_job_cnt.p, m.d_misc.p + 22);
                            size_t sb = 0;
                            cub::DeviceScan::ExclusiveSum(nullptr, sb, m.d_job_cnt.p, m.d_job_off.p, h_heavy + 1, st);
                            m.d_cub.ensure(sb + 16);
                            cub::DeviceScan::ExclusiveSum(m.d_cub.p, sb, m.d_job_cnt.p, m.d_job_off.p, h_heavy + 1, st);
                            uint32_t n_jobs = 0;
                            CK(cudaMemcpyAsync(&n_jobs, m.d_job_off.p + h_heavy, 4, cudaMemcpyDeviceToHost, st));
                            m.wait();
                            // a lane runs a whole job on its own (~5 ms of dependent rows): with only a few thousand jobs the warp-per-pair
                            // kernel, which spreads every Smith-Waterman over 32 lanes, is back sooner (measured on clean C2 pairs: 1 vs 5 ms)
                            if ((int)n_jobs < env_int("BSB_RESCUE_JOBS_MIN", 3000)) {
                                k_final_pe_heavy<<<heavy_blocks, 128, hv_smem, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_heavy);
                                m.launches += 3;
                                out.n_rescue_pairs = (uint64_t)h_heavy;
                            } else {
                            m.d_jobs.ensure((size_t)n_jobs + 1); m.d_job_res.ensure((size_t)n_jobs + 1);
                            k_rescue_enum<<<en_blocks, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, h_heavy, m.d_jobs.p, m.d_job_off.p, nullptr, m.d_misc.p + 23);
                            if (n_jobs) {
                                raise_dynamic_smem(k_rescue_sw, (int)rs_smem);
                                const int rs_bps = std::max(1, std::min(24, (int)((225 * 1024) / (rs_smem + 1024))));
                                const int rs_blocks = (int)std::min<size_t>(cdiv(n_jobs, 32), (size_t)I.n_sm * rs_bps);
                                const int cap_b = 128;
                                m.d_blist.ensure((size_t)rs_blocks * 32 * cap_b);
                                k_rescue_sw<<<rs_blocks, 32, rs_smem, st>>>(opt, I.ix, B, m.d_jobs.p, m.d_job_res.p, (int)n_jobs, rs_cells, m.d_blist.p, cap_b, m.d_misc.p + 24,
                                                                            env_int("BSB_RS_CHUNK", 32) & ~7);
                                ++m.launches;
                            }
                            const bool dbg_stats = getenv("BSB_DEBUG_STATS") != nullptr;
                            if (dbg_stats) CK(cudaMemsetAsync(m.d_work.p + 4, 0, 16, st));
                            const size_t rr_smem = ((size_t)2 * L.wreg_stride + L.reg_cap) * sizeof(AlnReg);   // the two lists of a pair + the sort's second buffer
                            if (rr_smem <= 220 * 1024 && !getenv("BSB_REPLAY_GLOBAL")) {
                                raise_dynamic_smem(k_rescue_replay<true>, (int)rr_smem);
                                const int rr_blocks = (int)std::min<size_t>((size_t)h_heavy, std::min<size_t>((size_t)fin_workers, (size_t)I.n_sm * std::max<size_t>(1, (220 * 1024) / (rr_smem + 1024))));
                                k_rescue_replay<true><<<rr_blocks, 32, rr_smem, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, h_heavy, m.d_jobs.p, m.d_job_res.p, m.d_job_off.p,
                                                                                      m.d_heavy2.p, m.d_misc.p + 13, m.d_misc.p + 25, dbg_stats ? m.d_work.p + 4 : nullptr);
                            } else
                            k_rescue_replay<false><<<en_blocks, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, h_heavy, m.d_jobs.p, m.d_job_res.p, m.d_job_off.p,
                                                                             m.d_heavy2.p, m.d_misc.p + 13, m.d_misc.p + 25, dbg_stats ? m.d_work.p + 4 : nullptr);
                            if (dbg_stats) {
                                unsigned long long sl[2] = {0, 0};
                                std::vector<uint32_t> cnt(h_heavy);
                                CK(cudaMemcpyAsync(sl, m.d_work.p + 4, 16, cudaMemcpyDeviceToHost, st));
                                CK(cudaMemcpyAsync(cnt.data(), m.d_job_cnt.p, (size_t)h_heavy * 4, cudaMemcpyDeviceToHost, st));
                                m.wait();
                                std::vector<int32_t> hv(h_heavy), nr(n);
                                CK(cudaMemcpy(hv.data(), heavy, (size_t)h_heavy * 4, cudaMemcpyDeviceToHost));
                                CK(cudaMemcpy(nr.data(), m.d_n_regs.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
                                const uint32_t k_slow = (uint32_t)sl[0];
                                const int p_slow = hv[k_slow];
                                fprintf(stderr, "[D::rescue] %d of %d pairs queued, %u Smith-Waterman jobs (max %u per pair); replay: slowest pair %d took %.2f M cycles of %.2f M in all "
                                        "(its jobs: %u, regions %d + %d)\n", h_heavy, n >> 1, n_jobs, *std::max_element(cnt.begin(), cnt.end()), p_slow, (double)(sl[0] >> 32) * 1024 / 1e6,
                                        (double)sl[1] * 1024 / 1e6, cnt[k_slow], nr[2 * p_slow], nr[2 * p_slow + 1]);
                            }
                            m.launches += 5;
                            k_final_pe_heavy<<<heavy_blocks, 128, hv_smem, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, m.d_heavy2.p, m.d_misc.p + 13, c_heavy);
                            ++m.launches;
                            out.n_rescue_jobs = n_jobs; out.n_rescue_pairs = (uint64_t)h_heavy;
                            }
                        }
                    } else {
                        k_final_pe_heavy<<<heavy_blocks, 128, hv_smem, st>>>(opt, I.ix, B, L, m.d_final_scratch.p, heavy, n_heavy, c_heavy);
                        ++m.launches;
                    }
                }
            }
            else k_final_se<<<fin_workers / fin_block, fin_block, 0, st>>>(opt, I.ix, B, L, m.d_final_scratch.p);
            ++m.launches;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&n_tasks, m.d_ntasks.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(m.ev[10], st));
        m.wait();
        T("select_done");
        if (n_tasks > m.task_cap) { m.task_cap = (size_t)n_tasks + (size_t)n_tasks / 4 + 4096; continue; }
        if (n_tasks) {
            const int slot_cap = 2 * max_q + 16, md_cap = 8 * max_q + 64, xb_cap = 4 * max_q + 64;
            const int dp_smem_per_warp = (2 * (max_q + 1) * 4 + max_q + 31) & ~15;
            const int fin_threads = I.n_sm * 8 * 128;
            m.d_task_cigar.ensure((size_t)n_tasks * slot_cap);
            m.d_task_ncig.ensure(n_tasks);
            m.d_task_text.ensure((size_t)fin_threads * (md_cap + xb_cap));
            k_tasks_dp<<<tk_blocks, tk_wpb * 32, tk_wpb * dp_smem_per_warp, st>>>(opt, I.ix, B, n_tasks, m.d_zbuf.p, z_cap, max_q, dp_smem_per_warp,
                                                                                   m.d_task_cigar.p, slot_cap, m.d_task_ncig.p, dyn_sched ? m.d_misc.p + 20 : nullptr);
            k_tasks_finish<<<fin_threads / 128, 128, 0, st>>>(opt, I.ix, B, n_tasks, m.d_task_cigar.p, slot_cap, m.d_task_ncig.p, m.d_task_text.p, md_cap, xb_cap,
                                                              dyn_sched ? m.d_misc.p + 21 : nullptr);
            m.launches += 2;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(m.ev[8], st));
        CK(cudaMemcpyAsync(&used, m.d_used.p, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out.reads.data(), m.d_out.p, (size_t)n * sizeof(ReadOut), cudaMemcpyDeviceToHost, st));
        m.wait();
        T("tasks_done");
        if (m.fin_scale < 64) {   // a data-dependent scratch limit was hit: more room, same stage again
            bool grow = false;
            for (int r = 0; r < n && !grow; ++r) grow = out.reads[r].err == ERR_SCRATCH_OVERFLOW;
            if (grow) {
                m.fin_scale *= 4;
                L = make_layout(m.fin_scale);
                m.d_final_scratch.ensure((size_t)std::max(fin_workers, heavy_blocks * 4) * L.total);
                if (used > m.arena_cap) m.arena_cap = (size_t)used + (size_t)used / 4 + (1 << 20);
                continue;
            }
        }
        if (used <= m.arena_cap) break;
        m.arena_cap = (size_t)used + (size_t)used / 4 + (1 << 20); // the counter keeps counting past the cap: exact retry size
    }
    // ---- SAM text on the device, D2H ----
    out.have_text = false; out.have_bam = false;
    size_t text_bytes = 0;
    if (dev_text && n) {
        SamView v;
        memset(&v, 0, sizeof v);
        v.flag = opt.flag; v.ch_conversion_threshold = opt.ch_conversion_threshold; v.ch_conversion_proportion = opt.ch_conversion_proportion;
        m.d_rg.ensure(out.rg_id.size() + 1);
        if (!out.rg_id.empty()) CK(cudaMemcpyAsync(m.d_rg.p, out.rg_id.data(), out.rg_id.size(), cudaMemcpyHostToDevice, st));
        v.rg_id = m.d_rg.p; v.rg_len = (int)out.rg_id.size();
        v.names = m.d_names.p; v.name_off = m.d_name_off.p; v.bases = B.bases; v.qual = m.d_qual.p; v.seq_off = B.seq_off;
        v.has_qual = m.d_has_qual.p; v.pattern = B.pattern; v.cmt = nullptr; v.cmt_off = nullptr;
        v.ctg_text = I.d_ctg_text.p; v.ctg_name_off = I.d_ctg_name_off.p; v.ctg_anno_off = I.d_ctg_anno_off.p;
        v.ctg_is_crick = I.d_ctg_is_crick.p; v.ctg_sign = I.d_ctg_sign.p;
        v.arena = m.d_arena.p; v.reads = m.d_out.p;
        m.d_text_len.ensure(n + 2); m.d_text_off.ensure(n + 2); m.d_stats.ensure(n + 1);
        CK(cudaMemsetAsync(m.d_text_len.p + n, 0, 4, st));
        k_sam_count<<<cdiv(n, 128), 128, 0, st>>>(v, n, pe ? 1 : 0, m.d_text_len.p, m.d_stats.p);
        size_t sb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, sb, m.d_text_len.p, m.d_text_off.p, n + 1, st);
        m.d_cub.ensure(sb + 16);
        cub::DeviceScan::ExclusiveSum(m.d_cub.p, sb, m.d_text_len.p, m.d_text_off.p, n + 1, st);
        uint32_t total = 0;
        CK(cudaMemcpyAsync(&total, m.d_text_off.p + n, 4, cudaMemcpyDeviceToHost, st));
        m.wait();
        text_bytes = total;
        m.d_text.ensure(text_bytes + 16);
        k_sam_write<<<cdiv(n, 128), 128, 0, st>>>(v, n, pe ? 1 : 0, m.d_text_off.p, m.d_text.p);
        m.launches += 3;
        CK(cudaGetLastError());
        CK(cudaEventRecord(m.ev[11], st));
        if (out.want_bam) {
            // ---- BAM: the arbiter, the records and their BGZF blocks are made here; only compressed blocks cross PCIe ----
            m.d_first.ensure(n + 1); m.d_rgrp.ensure(n + 1); m.d_bam_code.ensure(n + 1); m.d_bam_len.ensure(n + 2); m.d_bam_off.ensure(n + 2); m.d_bam_ctr.ensure(32);
            CK(cudaMemcpyAsync(m.d_first.p, b.first.data(), (size_t)n, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(m.d_rgrp.p, b.read_group.data(), (size_t)n, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(m.d_bam_ctr.p, 0, 16 * sizeof(unsigned long long), st));
            CK(cudaMemsetAsync(m.d_bam_len.p + n, 0, 4, st));
            BamView bv;
            bv.a.names = m.d_names.p; bv.a.name_off = m.d_name_off.p; bv.a.first = m.d_first.p; bv.a.read_group = m.d_rgrp.p; bv.a.stats = m.d_stats.p; bv.a.n = n;
            bv.text = m.d_text.p; bv.text_off = m.d_text_off.p; bv.code = m.d_bam_code.p;
            bv.ctg.text = I.d_ctg_text.p; bv.ctg.name_off = I.d_ctg_name_off.p; bv.ctg.sorted = I.d_ctg_sorted.p; bv.ctg.n = I.n_ctg;
            bv.bases = B.bases; bv.qual = m.d_qual.p; bv.seq_off = B.seq_off; bv.has_qual = m.d_has_qual.p;
            k_bam_arbiter<<<cdiv(n, 128), 128, 0, st>>>(bv.a, m.d_bam_code.p, m.d_bam_ctr.p);
            k_bam_count<<<cdiv(n, 128), 128, 0, st>>>(bv, m.d_bam_len.p, m.d_bam_ctr.p);
            cub::DeviceScan::ExclusiveSum(nullptr, sb, m.d_bam_len.p, m.d_bam_off.p, n + 1, st);
            m.d_cub.ensure(sb + 16);
            cub::DeviceScan::ExclusiveSum(m.d_cub.p, sb, m.d_bam_len.p, m.d_bam_off.p, n + 1, st);
            unsigned long long hc[16];
            uint32_t raw_total = 0;
            CK(cudaMemcpyAsync(hc, m.d_bam_ctr.p, sizeof hc, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&raw_total, m.d_bam_off.p + n, 4, cudaMemcpyDeviceToHost, st));
            m.wait();
            if (hc[8]) {
                const unsigned long long e = hc[8] - 1;
                throw std::runtime_error(std::string("[E::bam_encode] ") + bam_strerror((int)(e & 0xff)) + ": read '" + b.name((int)(e >> 8)) + "'");
            }
            const uint32_t max_entry = (uint32_t)hc[9];
            const int fixed = max_entry > 0xff00u / 2;
            const uint32_t quantum = bam_block_quantum(max_entry);
            const int nblk = raw_total ? (int)cdiv((size_t)raw_total, (size_t)quantum) + 1 : 0;
            uint32_t bgzf_total = 0;
            if (nblk) {
                m.d_bam_raw.ensure((size_t)raw_total + 64);
                k_bam_write<<<cdiv(n, 128), 128, 0, st>>>(bv, m.d_bam_off.p, reinterpret_cast<char *>(m.d_bam_raw.p));
                int df_per_sm;
                {   // (function attributes belong to the device: set once per aligner)
                    std::lock_guard<std::mutex> l(I.init_m);
                    if (!I.df_per_sm) {
                        CK(cudaFuncSetAttribute(k_bgzf_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateShared)));
                        int nb = 1;
                        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_bgzf_deflate, DF_CH, sizeof(DeflateShared)));
                        I.df_per_sm = std::max(1, std::min(nb, env_int("BSB_DF_PER_SM", nb)));   // fewer thread blocks per SM leave room for the other batches' kernels
                    }
                    df_per_sm = I.df_per_sm;
                }
                const int grid = std::min(nblk, I.n_sm * df_per_sm);
                m.d_bam_cut.ensure(nblk + 2); m.d_bgzf_len.ensure(nblk + 2); m.d_bgzf_off.ensure(nblk + 2);
                m.d_bgzf_slots.ensure((size_t)nblk * BGZF_SLOT); m.d_bam_tok.ensure((size_t)grid * (BGZF_MAX_IN + 8));
                k_bam_cuts<<<cdiv(nblk + 1, 128), 128, 0, st>>>(m.d_bam_off.p, n, quantum, fixed, nblk, m.d_bam_cut.p);
                CK(cudaMemsetAsync(m.d_bgzf_len.p + nblk, 0, 4, st));
                static const bool df_profile = getenv("BSB_DF_PROFILE") != nullptr;
                if (df_profile) CK(cudaMemsetAsync(m.d_bam_ctr.p + 16, 0, 16 * sizeof(unsigned long long), st));
                k_bgzf_deflate<<<grid, DF_CH, sizeof(DeflateShared), st>>>(m.d_bam_raw.p, m.d_bam_cut.p, nblk, m.d_bgzf_slots.p, m.d_bgzf_len.p, m.d_bam_tok.p,
                                                                         df_profile ? m.d_bam_ctr.p + 16 : nullptr);
                if (df_profile) {
                    unsigned long long pc[16];
                    CK(cudaMemcpyAsync(pc, m.d_bam_ctr.p + 16, sizeof pc, cudaMemcpyDeviceToHost, st));
                    m.wait();
                    unsigned long long tot = 0;
                    for (int k = 0; k < 12; ++k) tot += pc[k];
                    fprintf(stderr, "[D::deflate] %d blocks on %d thread blocks; cycles per phase (%%):", nblk, grid);
                    for (int k = 0; k < 12; ++k) fprintf(stderr, " %d:%.1f", k, 100.0 * (double)pc[k] / (double)(tot ? tot : 1));
                    fprintf(stderr, " | %.0f cycles per block\n", (double)tot / nblk);
                }
                cub::DeviceScan::ExclusiveSum(nullptr, sb, m.d_bgzf_len.p, m.d_bgzf_off.p, nblk + 1, st);
                m.d_cub.ensure(sb + 16);
                cub::DeviceScan::ExclusiveSum(m.d_cub.p, sb, m.d_bgzf_len.p, m.d_bgzf_off.p, nblk + 1, st);
                CK(cudaMemcpyAsync(&bgzf_total, m.d_bgzf_off.p + nblk, 4, cudaMemcpyDeviceToHost, st));
                m.wait();
                if (bgzf_total > (uint64_t)nblk * BGZF_SLOT) throw std::runtime_error("[E::bsbolt_b200] the BGZF stage produced an impossible size");
                m.d_bgzf_dense.ensure((size_t)bgzf_total + 64);
                k_bgzf_gather<<<nblk, 256, 0, st>>>(m.d_bgzf_slots.p, m.d_bgzf_len.p, m.d_bgzf_off.p, nblk, m.d_bgzf_dense.p);
                m.launches += 8;
                CK(cudaGetLastError());
            } else m.launches += 3;
            CK(cudaEventRecord(m.ev[12], st));
            out.bam.resize_uninit(bgzf_total);
            if (bgzf_total) CK(cudaMemcpyAsync(out.bam.data(), m.d_bgzf_dense.p, bgzf_total, cudaMemcpyDeviceToHost, st));
            for (int k = 0; k < 8; ++k) out.bam_counts[k] = hc[k];
            out.bam_raw_bytes = raw_total; out.bam_records = hc[10]; out.bam_blocks = (uint64_t)nblk;
            out.have_bam = true;
            text_bytes = bgzf_total;       // what crosses PCIe for this batch
        } else {
        out.text.resize_uninit(text_bytes); out.text_off.resize(n + 1); out.stats.resize(n);
        CK(cudaMemcpyAsync(out.text.data(), m.d_text.p, text_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out.text_off.data(), m.d_text_off.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out.stats.data(), m.d_stats.p, (size_t)n * sizeof(SamStats), cudaMemcpyDeviceToHost, st));
        }
        out.have_text = !out.have_bam;
    } else {
        out.arena.resize_uninit((size_t)used);
        CK(cudaMemcpyAsync(out.arena.data(), m.d_arena.p, (size_t)used, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(m.ev[9], st));
    m.wait();
    T("d2h_done");
    for (int r = 0; r < n; ++r)
        if (out.reads[r].err)
            throw std::runtime_error("[E::bsbolt_b200] read '" + b.name(r) + "' failed on the device with error code " + std::to_string(out.reads[r].err));
    float ms;
    for (int k = 0; k < 9; ++k) { CK(cudaEventElapsedTime(&ms, m.ev[k], m.ev[k + 1])); if (k < 8) out.ms_stage[k] = ms; else out.ms_d2h = ms; }
    out.ms_h2d = out.ms_stage[0];
    CK(cudaEventElapsedTime(&ms, m.ev[7], m.ev[10])); out.ms_select = ms;
    CK(cudaEventElapsedTime(&ms, m.ev[10], m.ev[8])); out.ms_tasks = ms;
    out.n_tasks = n_tasks;
    CK(cudaEventElapsedTime(&ms, m.ev[1], m.ev[8]));
    out.ms_kernels = ms;
    out.n_seeds = S;
    {
        unsigned long long w[4] = {0, 0, 0, 0};
        CK(cudaMemcpyAsync(w, m.d_work.p, sizeof w, cudaMemcpyDeviceToHost, st));
        m.wait();
        out.n_fm_ext = w[0]; out.n_fm_two_block = w[1]; out.n_ext_cells = w[2]; out.n_fm_two_block_ref = w[3];
        out.fm_block_bytes = I.ix.occ32 ? 32 : 64;
    }
    out.h2d_bytes = nb + (size_t)(n + 1) * 4 + (size_t)n;
    out.d2h_bytes = (out.have_bam ? text_bytes + 16 * sizeof(unsigned long long)
                     : out.have_text ? text_bytes + (size_t)(n + 1) * 4 + (size_t)n * sizeof(SamStats) : (size_t)used) + (size_t)n * sizeof(ReadOut);
    if (out.have_bam) { CK(cudaEventElapsedTime(&ms, m.ev[11], m.ev[12])); out.ms_bam = ms; }
    if (out.have_text || out.have_bam) {
        CK(cudaEventElapsedTime(&ms, m.ev[8], m.ev[11])); out.ms_text = ms;
        out.h2d_bytes += b.names.size() + (size_t)(n + 1) * 4 + nb + (size_t)n;
    }
}

// ------------------------------------------------------------------------------------------------
// measurement: random 32-byte sectors (the denominator of the seeding roofline)
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_random(uint4 *buf, uint64_t n16)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        buf[i] = make_uint4((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)(x * 3), (uint32_t)(x >> 17));
    }
}

// CHASE: the next sector index is a hash of the sector just read (a true dependency, like the rank that feeds the next
// bwt_extend); otherwise the indices follow a per-thread generator and UNROLL loads are in flight per thread.
template <bool CHASE>
__global__ void __launch_bounds__(256) k_random_sectors(const uint32_t *buf, uint64_t sector_mask, int steps, unsigned long long *sink)
{
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    if (CHASE) {
        for (int s = 0; s < steps; ++s) {
            const uint32_t *p = buf + ((x >> 11) & sector_mask) * 8;
            uint32_t v[8];
            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
            acc += v[0] ^ v[7];
            x = (x ^ ((uint64_t)v[3] << 32 | v[5])) * 0xBF58476D1CE4E5B9ull + s;
        }
    } else {
        for (int s = 0; s < steps; s += 4) {
            uint32_t v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                x = x * 6364136223846793005ull + 1442695040888963407ull;
                const uint32_t *p = buf + ((x >> 24) & sector_mask) * 8;
                asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]), "=r"(v[u][4]), "=r"(v[u][5]), "=r"(v[u][6]), "=r"(v[u][7]) : "l"(p));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u][0] ^ v[u][7];
        }
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

void random_sector_peak(int device, size_t footprint_bytes, double *gbs_independent, double *gbs_chase)
{
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    uint64_t bytes = 1ull << 26;                                         // the power of two nearest to the footprint asked for (64 MiB .. 16 GiB)
    while ((double)bytes * 1.4142 < (double)footprint_bytes && bytes < (16ull << 30)) bytes *= 2;
    const uint64_t n_sectors = bytes / 32;
    DevBuf<uint4> buf; buf.ensure(bytes / 16);
    DevBuf<unsigned long long> sink; sink.ensure(1);
    k_fill_random<<<prop.multiProcessorCount * 8, 256>>>(buf.p, bytes / 16);
    CK(cudaGetLastError());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, steps = 256;        // 2048 threads per SM
    double out[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {                              // first pass warms up; best of the next three
            CK(cudaEventRecord(e0));
            if (mode == 0) k_random_sectors<false><<<blocks, 256>>>(reinterpret_cast<const uint32_t *>(buf.p), n_sectors - 1, steps, sink.p);
            else k_random_sectors<true><<<blocks, 256>>>(reinterpret_cast<const uint32_t *>(buf.p), n_sectors - 1, steps, sink.p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        out[mode] = (double)blocks * 256 * steps * 32 / (best * 1e-3) / 1e9;
    }
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    if (gbs_independent) *gbs_independent = out[0];
    if (gbs_chase) *gbs_chase = out[1];
}

} // namespace bsb
