// bsb_hd.h -- host/device portability macros and the order-exact sorting primitive.
//
// Device functions are written once and compiled by nvcc for sm_100a (the product) and by g++
// for the CPU-side unit tests under tests/hostsim (test infrastructure only; the product library
// contains no CPU execution path for them).
// Attribution: introsort/comb_sort restate klib's ksort.h (ks_introsort, ks_combsort; MIT, Attractive Chaos) move for move --
// the order they leave equal keys in is visible in the reference's output. See NOTICE.md.
#pragma once
#include <stdint.h>
#include "bsb_types.h"

#if defined(__CUDACC__)
#define BSB_HD __host__ __device__ inline
#define BSB_HDN __host__ __device__ __noinline__
#else
#define BSB_HD inline
#define BSB_HDN inline
#endif

namespace bsb {

template <class T> BSB_HD T tmin(T a, T b) { return a < b ? a : b; }
template <class T> BSB_HD T tmax(T a, T b) { return a > b ? a : b; }
template <class T> BSB_HD void tswap(T &a, T &b) { T t = a; a = b; b = t; }
BSB_HD int iabs(int x) { return x < 0 ? -x : x; }

// 64-bit mix used for tie-breaking among equal-score hits (reference: utils.h:98-109).
BSB_HD uint64_t hash64(uint64_t key)
{
    key += ~(key << 32);
    key ^= (key >> 22);
    key += ~(key << 13);
    key ^= (key >> 8);
    key += (key << 3);
    key ^= (key >> 15);
    key += ~(key << 27);
    key ^= (key >> 31);
    return key;
}

// ---------------------------------------------------------------------------------------------
// Order-exact introsort.
//
// The reference sorts chains, alignment regions and pair candidates with klib's ks_introsort
// (ksort.h:176-226), which is NOT stable; with tied keys the output permutation is result-visible
// (which chain is extended first, which duplicate hit survives). The permutation is a pure
// function of the algorithm, so the same algorithm is restated here: median-of-three pivot moved
// to the right end, Hoare-style partition, sub-ranges of <=16 elements left for one final
// insertion sort, comb sort when the depth budget 2*ceil(log2 n) runs out.
// ---------------------------------------------------------------------------------------------
template <class T, class Lt>
BSB_HD void insertion_sort(T *s, T *t, Lt lt)
{
    for (T *i = s + 1; i < t; ++i)
        for (T *j = i; j > s && lt(*j, *(j - 1)); --j) tswap(*j, *(j - 1));
}

template <class T, class Lt>
BSB_HD void comb_sort(long n, T *a, Lt lt)
{
    const double shrink = 1.2473309501039786540366528676643;
    bool swapped;
    long gap = n;
    do {
        if (gap > 2) {
            gap = (long)(gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        swapped = false;
        for (T *i = a; i < a + n - gap; ++i) {
            T *j = i + gap;
            if (lt(*j, *i)) { tswap(*i, *j); swapped = true; }
        }
    } while (swapped || gap > 2);
    if (gap != 1) insertion_sort(a, a + n, lt);
}

template <class T, class Lt>
BSB_HD void introsort(long n, T *a, Lt lt)
{
    if (n < 1) return;
    if (n == 2) {
        if (lt(a[1], a[0])) tswap(a[0], a[1]);
        return;
    }
    struct Frame { T *l, *r; int depth; };
    Frame stack[48]; // only ranges of >16 elements are pushed and the smaller side is processed first: depth <= log2(n/16)
    Frame *top = stack;
    int d;
    for (d = 2; (1ul << d) < (unsigned long)n; ++d) {}
    d <<= 1;
    T *s = a, *t = a + (n - 1);
    for (;;) {
        if (s < t) {
            if (--d == 0) {
                comb_sort((long)(t - s) + 1, s, lt);
                t = s;
                continue;
            }
            T *i = s, *j = t, *k = i + ((j - i) >> 1) + 1;
            if (lt(*k, *i)) {
                if (lt(*k, *j)) k = j;
            } else k = lt(*j, *i) ? i : j;
            T rp = *k;
            if (k != t) tswap(*k, *t);
            for (;;) {
                do ++i; while (lt(*i, rp));
                do --j; while (i <= j && lt(rp, *j));
                if (j <= i) break;
                tswap(*i, *j);
            }
            tswap(*i, *t);
            if (i - s > t - i) {
                if (i - s > 16) { top->l = s; top->r = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->l = i + 1; top->r = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) {
                insertion_sort(a, a + n, lt);
                return;
            }
            --top; s = top->l; t = top->r; d = top->depth;
        }
    }
}

struct LtU64 { BSB_HD bool operator()(uint64_t a, uint64_t b) const { return a < b; } };

struct Pair64 { uint64_t x, y; };
struct LtPair64 {
    BSB_HD bool operator()(const Pair64 &a, const Pair64 &b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); }
};

} // namespace bsb
