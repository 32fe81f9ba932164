// bsb_chain.h -- seed chaining and chain filtering for one read (north_star stage 4).
//
//   chain_seeds()  <- mem_chain's seed loop + test_and_merge   (bwamem.c:194-215, 277-304)
//   chain_weight() <- mem_chain_weight                          (bwamem.c:217-236)
//   chain_filter() <- mem_chain_flt                             (bwamem.c:331-389)
//
// The reference keeps chains in a klib B-tree keyed by the chain's first reference position and
// merges each seed into its predecessor chain. With duplicate keys the predecessor that is found,
// and the in-order position of a new duplicate, depend on the tree's shape (kbtree.h:117-234), and
// the traversal order feeds an unstable sort -- so the tree itself is restated here: same order
// t=5 (KB_DEFAULT_SIZE 512 with a 40-byte key), same lower-bound search inside a node, same
// split-on-the-way-down insertion. Nodes hold chain indices and live in per-read scratch in HBM.
// Attribution: restates klib's kbtree.h (MIT, Attractive Chaos) and BWA-MEM's mem_chain / mem_chain_flt (bwamem.c, GPLv3,
// Heng Li): tree shape and unstable-sort order are result-visible. See NOTICE.md.
#pragma once
#include "bsb_index.h"

namespace bsb {

enum { BT_T = 5, BT_MAXK = 2 * BT_T - 1 };

struct BtNode {
    int32_t n, internal;
    int32_t key[BT_MAXK];
    int32_t child[BT_MAXK + 1];
};

struct ChainWS {           // per-read scratch, all sized from the read's seed count
    const Seed *seeds;     // in: the read's seeds in reference order
    int n_seeds;
    int32_t *next;         // [n_seeds] next seed of the same chain
    Chain *chains;         // [n_seeds] chain pool (insertion order)
    Chain *out;            // [n_seeds] chains in tree order, then filtered in place
    int32_t *tmp;          // [n_seeds] kept-list scratch
    BtNode *nodes;         // [n_seeds/4 + 2]
    int node_cap;
};

struct ChainTree {
    BtNode *nodes; int n_nodes, cap, root, n_keys;
    const Chain *chains;
    int *err;

    BSB_HD int cmp_pos(int64_t a, int64_t b) const { return (b < a) - (a < b); }

    BSB_HD int new_node(int internal)
    {
        if (n_nodes >= cap) { *err = ERR_SCRATCH_OVERFLOW; return 0; }
        BtNode &z = nodes[n_nodes];
        z.n = 0; z.internal = internal;
        return n_nodes++;
    }
    BSB_HD void init() { n_nodes = 0; n_keys = 0; root = new_node(0); }

    // index of the last key <= k inside node x (first of equals), r = comparison with that key
    BSB_HD int find(const BtNode &x, int64_t kpos, int *r) const
    {
        int begin = 0, end = x.n;
        if (x.n == 0) return -1;
        while (begin < end) {
            int mid = (begin + end) >> 1;
            if (cmp_pos(chains[x.key[mid]].pos, kpos) < 0) begin = mid + 1;
            else end = mid;
        }
        if (begin == x.n) { *r = 1; return x.n - 1; }
        if ((*r = cmp_pos(kpos, chains[x.key[begin]].pos)) < 0) --begin;
        return begin;
    }

    // closest chain at or before kpos, -1 if none
    BSB_HD int lower(int64_t kpos) const
    {
        int lo = -1, x = root, r = 0;
        for (;;) {
            const BtNode &nd = nodes[x];
            int i = find(nd, kpos, &r);
            if (i >= 0 && r == 0) return nd.key[i];
            if (i >= 0) lo = nd.key[i];
            if (!nd.internal) return lo;
            x = nd.child[i + 1];
        }
    }

    BSB_HD void split(int xi, int i, int yi)
    {
        int zi = new_node(nodes[yi].internal);
        BtNode &x = nodes[xi], &y = nodes[yi], &z = nodes[zi];
        z.n = BT_T - 1;
        for (int j = 0; j < BT_T - 1; ++j) z.key[j] = y.key[BT_T + j];
        if (y.internal) for (int j = 0; j < BT_T; ++j) z.child[j] = y.child[BT_T + j];
        y.n = BT_T - 1;
        for (int j = x.n; j > i; --j) x.child[j + 1] = x.child[j];
        x.child[i + 1] = zi;
        for (int j = x.n - 1; j >= i; --j) x.key[j + 1] = x.key[j];
        x.key[i] = y.key[BT_T - 1];
        ++x.n;
    }

    BSB_HD void put(int ci)
    {
        int64_t kpos = chains[ci].pos;
        ++n_keys;
        int r = root, dummy;
        if (nodes[r].n == BT_MAXK) {
            int s = new_node(1);
            nodes[s].child[0] = r;
            root = s;
            split(s, 0, r);
            r = s;
        }
        int x = r;
        for (;;) {
            BtNode &nd = nodes[x];
            if (!nd.internal) {
                int i = find(nd, kpos, &dummy);
                for (int j = nd.n - 1; j > i; --j) nd.key[j + 1] = nd.key[j];
                nd.key[i + 1] = ci;
                ++nd.n;
                return;
            }
            int i = find(nd, kpos, &dummy) + 1;
            if (nodes[nd.child[i]].n == BT_MAXK) {
                split(x, i, nd.child[i]);
                if (cmp_pos(kpos, chains[nodes[x].key[i]].pos) > 0) ++i;
            }
            x = nodes[x].child[i];
        }
    }

    // in-order traversal into out[]; returns the number of chains
    BSB_HD int traverse(const Chain *pool, Chain *out) const
    {
        int sx[16], si[16], sp = 0, n = 0;
        sx[0] = root; si[0] = 0;
        for (;;) {
            // descend
            while (sx[sp] >= 0 && si[sp] <= nodes[sx[sp]].n) {
                const BtNode &nd = nodes[sx[sp]];
                int c = nd.internal ? nd.child[si[sp]] : -1;
                ++sp; sx[sp] = c; si[sp] = 0;
            }
            --sp;
            if (sp < 0) break;
            if (sx[sp] >= 0 && si[sp] < nodes[sx[sp]].n) out[n++] = pool[nodes[sx[sp]].key[si[sp]]];
            ++si[sp];
        }
        return n;
    }
};

// returns 1 if the seed was absorbed by chain c
BSB_HD int chain_try_merge(const Opt &opt, int64_t l_pac, Chain &c, const Seed *seeds, int32_t *next, int si)
{
    const Seed &p = seeds[si];
    const Seed &first = seeds[c.head], &last = seeds[c.tail];
    int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
    if (p.rid != c.rid) return 0;
    if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend)
        return 1; // contained
    if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0; // other strand
    int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
    if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - last.len < opt.max_chain_gap && y - last.len < opt.max_chain_gap) {
        next[c.tail] = si; next[si] = -1;
        c.tail = si; ++c.n;
        return 1;
    }
    return 0;
}

BSB_HD int chain_weight(const Chain &c, const Seed *seeds, const int32_t *next)
{
    int64_t end = 0;
    int w = 0, tmp;
    for (int j = c.head; j >= 0; j = next[j]) {
        const Seed &s = seeds[j];
        if (s.qbeg >= end) w += s.len;
        else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
        end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
    }
    tmp = w; w = 0; end = 0;
    for (int j = c.head; j >= 0; j = next[j]) {
        const Seed &s = seeds[j];
        if (s.rbeg >= end) w += s.len;
        else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
        end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
    }
    w = w < tmp ? w : tmp;
    return w < (1 << 30) ? w : (1 << 30) - 1;
}

struct LtChainW { BSB_HD bool operator()(const Chain &a, const Chain &b) const { return a.w > b.w; } };

// Builds chains from the read's seeds (already in reference iteration order) and returns the
// number of chains left in ws.out[] (tree order).
BSB_HD int chain_seeds(const Opt &opt, const IndexView &ix, ChainWS &ws, int l_rep, int l_seq, int *err)
{
    ChainTree tree;
    tree.nodes = ws.nodes; tree.cap = ws.node_cap; tree.chains = ws.chains; tree.err = err;
    tree.init();
    int n_pool = 0;
    for (int si = 0; si < ws.n_seeds; ++si) {
        const Seed &s = ws.seeds[si];
        if (s.rid < 0) continue;
        int to_add = 0;
        if (tree.n_keys) {
            int lo = tree.lower(s.rbeg);
            if (lo < 0 || !chain_try_merge(opt, ix.l_pac, ws.chains[lo], ws.seeds, ws.next, si)) to_add = 1;
        } else to_add = 1;
        if (to_add) {
            Chain &c = ws.chains[n_pool];
            c.pos = s.rbeg; c.n = 1; c.head = c.tail = si; c.rid = s.rid; c.first = -1;
            c.w = 0; c.kept = 0; c.is_alt = (int8_t)(ix.anns[s.rid].is_alt != 0); c.frac_rep = 0;
            ws.next[si] = -1;
            tree.put(n_pool);
            ++n_pool;
            if (*err) return 0;
        }
    }
    int n = tree.traverse(ws.chains, ws.out);
    float fr = (float)l_rep / l_seq;
    for (int i = 0; i < n; ++i) ws.out[i].frac_rep = fr;
    return n;
}

BSB_HD int chn_beg(const Chain &c, const Seed *seeds) { return seeds[c.head].qbeg; }
BSB_HD int chn_end(const Chain &c, const Seed *seeds) { return seeds[c.tail].qbeg + seeds[c.tail].len; }

// filters ws.out[0..n) in place; returns the number kept
BSB_HD int chain_filter(const Opt &opt, ChainWS &ws, int n_chn)
{
    Chain *a = ws.out;
    const Seed *sd = ws.seeds;
    int i, k;
    if (n_chn == 0) return 0;
    for (i = k = 0; i < n_chn; ++i) {
        Chain &c = a[i];
        c.first = -1; c.kept = 0;
        c.w = chain_weight(c, sd, ws.next);
        if (c.w < opt.min_chain_weight) continue;
        a[k++] = c;
    }
    n_chn = k;
    if (n_chn == 0) return 0; // (the reference would touch a[0] here; only reachable with -W > 0)
    introsort((long)n_chn, a, LtChainW());
    int *kept = ws.tmp, n_kept = 0;
    a[0].kept = 3;
    kept[n_kept++] = 0;
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        for (k = 0; k < n_kept; ++k) {
            int j = kept[k];
            int b_max = tmax(chn_beg(a[j], sd), chn_beg(a[i], sd));
            int e_min = tmin(chn_end(a[j], sd), chn_end(a[i], sd));
            if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
                int li = chn_end(a[i], sd) - chn_beg(a[i], sd);
                int lj = chn_end(a[j], sd) - chn_beg(a[j], sd);
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
                    large_ovlp = 1;
                    if (a[j].first < 0) a[j].first = i;
                    if (a[i].w < a[j].w * opt.drop_ratio && a[j].w - a[i].w >= opt.min_seed_len << 1) break;
                }
            }
        }
        if (k == n_kept) {
            kept[n_kept++] = i;
            a[i].kept = large_ovlp ? 2 : 3;
        }
    }
    for (i = 0; i < n_kept; ++i) {
        Chain &c = a[kept[i]];
        if (c.first >= 0) a[c.first].kept = 1;
    }
    for (i = k = 0; i < n_chn; ++i) {
        if (a[i].kept == 0 || a[i].kept == 3) continue;
        if (++k >= opt.max_chain_extend) break;
    }
    for (; i < n_chn; ++i)
        if (a[i].kept < 3) a[i].kept = 0;
    for (i = k = 0; i < n_chn; ++i)
        if (a[i].kept != 0) a[k++] = a[i];
    return k;
}

} // namespace bsb
