// bsb_stages.h -- the per-read / per-seed / per-pair work items of one batch, expressed over flat
// HBM buffers. Each stage_* function is the body of one kernel (kernels.cu launches them over the
// batch); tests/hostsim runs the same bodies in a CPU loop to unit-test the logic without a GPU.
//
// Batch buffers (device pointers in the product):
//   per base  : bases (ASCII in), seq (converted codes), oseq (original codes)
//   per read  : seq_off[n+1], pattern, intervals [r*intv_cap ..), n_intv, l_rep, n_seed,
//               seed_off[n+1] (exclusive scan of n_seed), n_chain, n_regs, err
//   per seed  : seeds, next, chain pool, chains, tmp, cseeds, srt, regs  (all indexed seed_off[r] + j)
//   per worker: scratch blocks (interval lists, DP rows, traceback, strings)
#pragma once
#include "bsb_seed3.h"
#include "bsb_chain.h"
#include "bsb_extend.h"
#include "bsb_final.h"

namespace bsb {

struct BatchDev {
    int n;                       // bseq entries
    int is_pe;
    int64_t n_processed;
    const char *bases;
    const uint32_t *seq_off;
    const uint8_t *pattern;
    uint8_t *seq, *oseq;
    // seeding
    Intv *intv; int intv_cap;
    int32_t *n_intv, *l_rep, *n_seed;
    // seeds and everything sized by them
    const uint32_t *seed_off;    // n+1
    Seed *seeds;
    int32_t *next;
    Chain *chain_pool, *chains;
    int32_t *tmp;
    Seed *cseeds;
    uint64_t *srt;
    AlnReg *regs;
    BtNode *nodes;
    int32_t *n_chain, *n_regs;
    int32_t *err;
    // pairing
    int8_t *pe_dir; int64_t *pe_isize;
    PeStat pes[4];
    MathTab mt;
    // output
    ReadOut *out;
    Arena arena;
    TaskList tasks;
};


BSB_HD uint8_t nt4_code(unsigned char c)
{   // nst_nt4_table (bntseq.c:48-65)
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

// K1: in-silico bisulfite conversion of one base (bsConversion acts on upper-case C/G only)
BSB_HD void stage_convert_base(const BatchDev &B, int r, uint32_t i)
{
    unsigned char c = (unsigned char)B.bases[i];
    B.oseq[i] = nt4_code(c);
    if (B.pattern[r]) { if (c == 'G') c = 'A'; }
    else { if (c == 'C') c = 'T'; }
    uint8_t code = nt4_code(c);
    B.seq[i] = code == 5 ? 4 : code; // mem_align1_core maps through nst_nt4_table where '-' is 5; any code > 3 is "ambiguous"
}

// tail of K2 for one read: frac_rep bookkeeping + seed count (bwamem.c:269-283) from the sorted interval list
BSB_HD void seed_finish(const Opt &opt, const BatchDev &B, int r, const Intv *mem, int n, int err)
{
    if (err) { B.err[r] = err; return; }
    int b = 0, e = 0, l_rep = 0, total = 0;
    for (int i = 0; i < n; ++i) {
        const Intv &p = mem[i];
        int sb = (int)(p.info >> 32), se = (int)(uint32_t)p.info;
        int64_t step = p.x2 > (uint64_t)opt.max_occ ? (int64_t)(p.x2 / opt.max_occ) : 1;
        int64_t cnt = ((int64_t)p.x2 + step - 1) / step;
        total += (int)(cnt < opt.max_occ ? cnt : opt.max_occ);
        if (p.x2 <= (uint64_t)opt.max_occ) continue;
        if (sb > e) { l_rep += e - b; b = sb; e = se; }
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    B.n_intv[r] = n; B.l_rep[r] = l_rep; B.n_seed[r] = total;
}

// K3: one suffix-array lookup. g = global seed slot, r = owning read (seed_off[r] <= g < seed_off[r+1])
BSB_HD void stage_sa(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, uint32_t g)
{
    uint32_t j = g - B.seed_off[r];
    const Intv *mem = B.intv + (size_t)r * B.intv_cap;
    int n = B.n_intv[r];
    for (int i = 0; i < n; ++i) {
        const Intv &p = mem[i];
        int64_t step = 1, cnt = (int64_t)p.x2;      // bwamem.c:282-283; the 64-bit divisions only for repetitive intervals
        if (p.x2 > (uint64_t)opt.max_occ) {
            step = (int64_t)(p.x2 / opt.max_occ);
            cnt = ((int64_t)p.x2 + step - 1) / step;
            if (cnt > opt.max_occ) cnt = opt.max_occ;
        }
        if (j < (uint32_t)cnt) {
            Seed s;
            int slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
            s.rbeg = (int64_t)fm_sa(ix, p.x0 + (uint64_t)(j * step));
            s.qbeg = (int32_t)(p.info >> 32);
            s.score = s.len = slen;
            s.rid = intv2rid(ix, s.rbeg, s.rbeg + s.len);
            B.seeds[g] = s;
            return;
        }
        j -= (uint32_t)cnt;
    }
}

// K4: chaining + chain filter for read r; leaves B.chains/B.cseeds compacted, B.n_chain[r] set
BSB_HD void stage_chain(const Opt &opt, const IndexView &ix, const BatchDev &B, int r)
{
    const uint32_t so = B.seed_off[r];
    const int ns = (int)(B.seed_off[r + 1] - so);
    B.n_chain[r] = 0;
    if (ns == 0 || B.err[r]) return;
    ChainWS ws;
    ws.seeds = B.seeds + so; ws.n_seeds = ns; ws.next = B.next + so;
    ws.chains = B.chain_pool + so; ws.out = B.chains + so; ws.tmp = B.tmp + so;
    ws.nodes = B.nodes + ((so >> 2) + 2 * (size_t)r);
    ws.node_cap = (ns >> 2) + 2;
    int err = 0;
    const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    int n = chain_seeds(opt, ix, ws, B.l_rep[r], len, &err);
    if (err) { B.err[r] = err; return; }
    n = chain_filter(opt, ws, n);
    // compact each chain's seeds into cseeds, in chain order
    Seed *cs = B.cseeds + so;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        Chain &c = ws.out[i];
        int head = k;
        for (int j = c.head; j >= 0; j = ws.next[j]) cs[k++] = ws.seeds[j];
        c.head = head; c.tail = k - 1;
    }
    B.n_chain[r] = n;
}

// K4b: mem_flt_chained_seeds for read r (only reads of about 720 bp and more, or any read when -W is set). Scratch: four
// rows of SEED_SW_CAP cells (the flanked window is shorter than 200 bases, padded to a multiple of 8).
enum { SEED_SW_CAP = 208 };
BSB_HD void stage_seed_sw(const Opt &opt, const IndexView &ix, const BatchDev &B, int r, const double *log_tab, int n_log)
{
    const uint32_t so = B.seed_off[r];
    const int nc = B.n_chain[r];
    if (nc == 0 || B.err[r]) return;
    const int len = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    int err = 0;
    const int min_hsp = seed_sw_min_score(opt, len, log_tab, n_log, &err);
    if (err) { B.err[r] = err; return; }
    if (min_hsp < 0) return;
    int32_t rows[4 * SEED_SW_CAP];
    SwScratch ws = {rows, rows + SEED_SW_CAP, rows + 2 * SEED_SW_CAP, rows + 3 * SEED_SW_CAP, nullptr, SEED_SW_CAP, 0};
    filter_chained_seeds(opt, ix, len, B.seq + B.seq_off[r], nc, B.chains + so, B.cseeds + so, min_hsp, ws, &err);
    if (err) B.err[r] = err;
}

// PE: insert-size candidate of pair p
BSB_HD void stage_pestat(const Opt &opt, const IndexView &ix, const BatchDev &B, int p)
{
    int r0 = p << 1, r1 = r0 | 1;
    int64_t is = 0;
    int d = pestat_candidate(opt, ix.l_pac, B.regs + B.seed_off[r0], B.n_regs[r0], B.regs + B.seed_off[r1], B.n_regs[r1], &is);
    B.pe_dir[p] = (int8_t)d; B.pe_isize[p] = is;
}

BSB_HD void readout_init(ReadOut &ro, int err)
{
    ro.aln_off = 0; ro.n_aln = 0; ro.h_pos = -1; ro.h_rid = -1; ro.h_is_rev = 0; ro.h_n_cigar = 0; ro.h_rlen = 0;
    ro.h_ch_meth = ro.h_ch_unmeth = 0; ro.err = err; ro.pad_ = 0;
}

// K8a (single-end): mark primary, choose the records, MAPQ/flags, XA lists; queues the alignments (AlnTask).
// Works on a private copy of the regions (wregs) so that a batch can be re-finalised after an
// arena overflow: mark_primary's hash tie-break depends on the incoming order.
BSB_HD void stage_final_se(const Opt &opt, const IndexView &ix, BatchDev &B, int r, FinalWS &ws, AlnReg *wregs)
{
    ReadOut &ro = B.out[r];
    readout_init(ro, B.err[r]);
    if (ro.err) return;
    int err = 0;
    const int n = B.n_regs[r];
    if (n > ws.reg_cap) { ro.err = ERR_SCRATCH_OVERFLOW; return; }
    const AlnReg *src = B.regs + B.seed_off[r];
    for (int j = 0; j < n; ++j) wregs[j] = src[j];
    mark_primary(opt, n, wregs, B.n_processed + r, ws.z);
    if (opt.flag & F_PRIMARY5) reorder_primary5(opt.T, n, wregs);
    ReadCtx rc;
    rc.read = r;
    rc.l_seq = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    rc.regs = wregs; rc.n_regs = n;
    emit_read(opt, B.mt, rc, 0, -1, ws, B.arena, B.tasks, B.out, &err);
    ro.err = err;
}

// K7 + K8a (paired-end): mate rescue, pairing, record selection. wregs: per-worker scratch of
// 2*(reg_cap + max_matesw) regions
// `defer` (optional): returns true, with nothing written, when the pair needs rescue Smith-Waterman (see mate_rescue)
// `sink`: only enumerates the rescue Smith-Waterman jobs of the pair; `pre`: finalises the pair with their results (bsb_final.h)
BSB_HD bool stage_final_pe(const Opt &opt, const IndexView &ix, BatchDev &B, int p, FinalWS &ws, AlnReg *wregs, bool defer = false,
                           RescueSink *sink = nullptr, const RescuePre *pre = nullptr)
{
    const int r0 = p << 1, r1 = r0 | 1;
    const int e0 = B.err[r0] ? B.err[r0] : B.err[r1];
    if (!sink) { readout_init(B.out[r0], e0); readout_init(B.out[r1], e0); }
    if (e0) return false;
    const int stride = ws.reg_cap + opt.max_matesw;
    RegList rl[2];
    for (int i = 0; i < 2; ++i) {
        int r = r0 | i, n = B.n_regs[r];
        rl[i].a = wregs + (size_t)i * stride; rl[i].cap = ws.reg_cap; rl[i].n = n;
        if (n > ws.reg_cap) { B.out[r0].err = B.out[r1].err = ERR_SCRATCH_OVERFLOW; return false; }
        const AlnReg *src = B.regs + B.seed_off[r];
        for (int j = 0; j < n; ++j) rl[i].a[j] = src[j];
    }
    int err = 0, deferred = 0;
    finalize_pair(opt, ix, B.mt, B.pes, (uint64_t)((B.n_processed >> 1) + p), r0,
                  (int)(B.seq_off[r0 + 1] - B.seq_off[r0]), B.seq + B.seq_off[r0], rl[0],
                  (int)(B.seq_off[r1 + 1] - B.seq_off[r1]), B.seq + B.seq_off[r1], rl[1],
                  ws, B.arena, B.tasks, B.out, &err, defer ? &deferred : nullptr, sink, pre);
    if (deferred) return true;
    if (sink) return false;
    if (err) B.out[r0].err = B.out[r1].err = err;
    return false;
}

// K6 + K8b (scalar form): one queued alignment -> CIGAR, NM, MD/XB, position, written to its record(s)
BSB_HD void stage_task(const Opt &opt, const IndexView &ix, BatchDev &B, unsigned int k, FinalWS &ws)
{
    const AlnTask t = B.tasks.a[k];
    const int r = t.read;
    const int l = (int)(B.seq_off[r + 1] - B.seq_off[r]);
    AlnBody b;
    int err = 0;
    if (align_body(opt, ix, t, l, B.seq + B.seq_off[r], B.oseq + B.seq_off[r], ws, b, &err))
        task_store(ix, t, b, ws.cigar, ws.md, B.arena, B.out, &err);
    if (err) B.out[r].err = err;
}

} // namespace bsb
