// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Runs the kernel bodies of bsbolt_b200/csrc/bsb_stages.h in plain CPU loops so that the alignment
// logic can be unit-tested against the reference (oracle/_ref) in the GPU-less build container.
// The product path is kernels.cu/pipeline.cu; it has no CPU execution mode.
//
// Usage: hostsim mem [bwa-mem options] <idxbase> <in1.fq> [in2.fq]   (SAM on stdout)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <stdexcept>
#include <vector>
#include <mutex>
#include "../../bsbolt_b200/csrc/bsb_stages.h"
#include "scalar_stages.h"
#include "../../bsbolt_b200/csrc/bsb_extlane.h"
#include "../../bsbolt_b200/csrc/bsb_rescue.h"
#include "../../bsbolt_b200/csrc/host_mem.h"
#include "../../bsbolt_b200/csrc/host_bam.h"
#include "bam_emulation.h"

using namespace bsb;

class HostSimAligner : public BatchAligner {
public:
    explicit HostSimAligner(const HostIndex &idx) : idx_(idx), ix_(idx.host_view())
    {
        build_log_table(log_tab_, 65536);
        if (getenv("BSB_HOSTSIM_SEED_V3") && !getenv("BSB_HOSTSIM_REF_BLOCKS") && idx.seq_len + 1 < (1ull << 32)) {
            const uint64_t nb32 = (uint64_t)(idx.bwt.size() / 16) * 2;   // the product's sector-sized occ blocks
            occ32_.resize(nb32 * 8);
            for (uint64_t b = 0; b < nb32; ++b) occ32_make_block(idx.bwt.data(), b, occ32_.data() + b * 8);
            ix_.occ32 = occ32_.data();
        }
    }

    void align(const Opt &opt, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int = 0) override
    {
        const int n = b.n;
        const bool pe = (opt.flag & F_PE) != 0;
        int max_len = 0;
        for (int i = 0; i < n; ++i) max_len = std::max(max_len, b.len(i));
        BatchDev B;
        memset(&B, 0, sizeof B);
        B.n = n; B.is_pe = pe; B.n_processed = n_processed;
        B.bases = b.bases.data(); B.seq_off = b.seq_off.data(); B.pattern = b.pattern.data();
        std::vector<uint8_t> seq(b.bases.size() + 1), oseq(b.bases.size() + 1);
        B.seq = seq.data(); B.oseq = oseq.data();
        B.intv_cap = std::max(256, 2 * max_len);
        std::vector<Intv> intv, sc;
        std::vector<int32_t> n_intv(n), l_rep(n), n_seed(n), n_chain(n), n_regs(n), err(n, 0);
        B.n_intv = n_intv.data(); B.l_rep = l_rep.data(); B.n_seed = n_seed.data();
        B.n_chain = n_chain.data(); B.n_regs = n_regs.data(); B.err = err.data();
        for (int r = 0; r < n; ++r)
            for (uint32_t i = b.seq_off[r]; i < b.seq_off[r + 1]; ++i) stage_convert_base(B, r, i);
        for (;;) { // an interval-list overflow is retried with twice the capacity, never truncated
            intv.assign((size_t)n * B.intv_cap, Intv()); sc.assign((size_t)3 * B.intv_cap, Intv());
            B.intv = intv.data();
            SeedScratch ss = {sc.data(), sc.data() + B.intv_cap, sc.data() + 2 * B.intv_cap};
            bool ovf = false;
            const bool use_sm = getenv("BSB_HOSTSIM_SEED_SM") != nullptr;
            const bool use_v3 = getenv("BSB_HOSTSIM_SEED_V3") != nullptr;
            for (int r = 0; r < n; ++r) { err[r] = 0; stage_seed(opt, ix_, B, r, ss, use_sm, true, use_v3); if (err[r] == ERR_INTV_OVERFLOW) ovf = true; }
            if (!ovf) break;
            B.intv_cap *= 2;
        }
        std::vector<uint32_t> seed_off(n + 1, 0);
        for (int r = 0; r < n; ++r) seed_off[r + 1] = seed_off[r] + (uint32_t)n_seed[r];
        const size_t S = seed_off[n];
        B.seed_off = seed_off.data();
        std::vector<Seed> seeds(S + 1), cseeds(S + 1);
        std::vector<int32_t> next(S + 1), tmp(S + 1);
        std::vector<Chain> pool(S + 1), chains(S + 1);
        std::vector<uint64_t> srt(S + 1);
        std::vector<AlnReg> regs(S + 1);
        std::vector<BtNode> nodes(S / 4 + 2 * (size_t)n + 4);
        B.seeds = seeds.data(); B.cseeds = cseeds.data(); B.next = next.data(); B.tmp = tmp.data();
        B.chain_pool = pool.data(); B.chains = chains.data(); B.srt = srt.data(); B.regs = regs.data(); B.nodes = nodes.data();
        for (int r = 0; r < n; ++r)
            for (uint32_t g = seed_off[r]; g < seed_off[r + 1]; ++g) stage_sa(opt, ix_, B, r, g);
        for (int r = 0; r < n; ++r) stage_chain(opt, ix_, B, r);
        for (int r = 0; r < n; ++r) stage_seed_sw(opt, ix_, B, r, log_tab_.data(), (int)log_tab_.size());
        // DP scratch
        const int max_q = max_len + 8;
        std::vector<int32_t> eh(2 * (size_t)(max_q + 1));
        const long z_cap = (long)(max_q) * (long)(max_q + 2 * (4 * opt.w) + 64);
        std::vector<uint8_t> z(z_cap);
        DpScratch dp = {eh.data(), z.data(), z_cap, max_q};
        if (getenv("HOSTSIM_EXT_LANES")) {
            // the per-read extension machine of k_extend_lanes (bsb_extlane.h), LANES reads interleaved row by row like the
            // lanes of a warp: control flow up to the next DP row, then one row each; then the tail, read by read
            const int LANES = std::max(1, atoi(getenv("HOSTSIM_EXT_LANES")));
            const int chunk = getenv("HOSTSIM_EXT_CHUNK") ? std::max(1, atoi(getenv("HOSTSIM_EXT_CHUNK"))) : 16;
            std::vector<std::vector<uint32_t>> rows(LANES, std::vector<uint32_t>(max_q + 2));
            std::vector<ExtLane<PackedRow<1>>> L(LANES);
            std::vector<int> tg(max_q + 2), twl(max_q + 2), twr(max_q + 2);
            const int amax = ext_amax(opt);
            for (int q = 0; q < max_q + 2; ++q) ext_tables_fill(opt, amax, q, tg.data(), twl.data(), twr.data());
            const ExtTables tabs = {tg.data(), twl.data(), twr.data(), getenv("HOSTSIM_EXT_NOTAB") ? 0 : max_q + 2, amax};
            for (int l = 0; l < LANES; ++l) { L[l].state = ExtLane<PackedRow<1>>::IDLE; L[l].H.p = rows[l].data(); L[l].row_cap = max_q + 2; L[l].T = tabs; L[l].n_cells = 0; }
            int next = 0;
            for (;;) {
                bool any = false;
                for (int l = 0; l < LANES; ++l) {
                    ExtLane<PackedRow<1>> &m = L[l];
                    while (m.state == ExtLane<PackedRow<1>>::IDLE && next < n) { m.begin_read(B, next++); m.advance(opt, ix_, B, max_q); }
                    if (m.state != ExtLane<PackedRow<1>>::IDLE && m.state != ExtLane<PackedRow<1>>::ROW) m.advance(opt, ix_, B, max_q);
                    while (m.state == ExtLane<PackedRow<1>>::IDLE && next < n) { m.begin_read(B, next++); m.advance(opt, ix_, B, max_q); }
                    if (m.state == ExtLane<PackedRow<1>>::ROW) { m.step(opt, ix_, chunk); any = true; }
                }
                if (!any && next >= n) {
                    bool busy = false;
                    for (int l = 0; l < LANES; ++l) busy = busy || L[l].state != ExtLane<PackedRow<1>>::IDLE;
                    if (!busy) break;
                }
            }
            for (int r = 0; r < n; ++r) extend_tail(opt, ix_, B, r, dp);
        } else
        for (int r = 0; r < n; ++r) stage_extend(opt, ix_, B, r, dp);
        int max_regs = 0;
        for (int r = 0; r < n; ++r) max_regs = std::max(max_regs, n_regs[r]);
        // pairing statistics
        std::vector<double> pair_tab;
        if (pe) {
            std::vector<int8_t> dir(n / 2);
            std::vector<int64_t> isz(n / 2);
            B.pe_dir = dir.data(); B.pe_isize = isz.data();
            if (pes0) memcpy(B.pes, pes0, sizeof B.pes);
            else {
                for (int p = 0; p < n / 2; ++p) stage_pestat(opt, ix_, B, p);
                estimate_pestat(opt, dir, isz, B.pes, 3);
            }
            memcpy(out.pes, B.pes, sizeof B.pes);
        }
        build_pair_table(opt, B.pes, pair_tab, B.mt.pair_off);
        B.mt.pair_tab = pair_tab.data();
        B.mt.log_tab = log_tab_.data(); B.mt.n_log = (int)log_tab_.size();
        // finalisation scratch
        FinalWS ws;
        ws.dp = dp;
        const int reg_cap = max_regs + 4 * opt.max_matesw + 8;
        std::vector<uint32_t> cigar(2 * max_q + 16);
        std::vector<char> md(8 * max_q + 64), xb(4 * max_q + 64);
        std::vector<int32_t> cnt(reg_cap), zz(reg_cap);
        std::vector<int8_t> has_alt(reg_cap);
        const int pair_cap = 16384;
        std::vector<Pair64> pv(pair_cap), pu(pair_cap);
        ws.cigar = cigar.data(); ws.cigar_cap = (int)cigar.size();
        ws.md = md.data(); ws.md_cap = (int)md.size(); ws.xb = xb.data(); ws.xb_cap = (int)xb.size();
        ws.cnt = cnt.data(); ws.has_alt = has_alt.data(); ws.z = zz.data();
        ws.pv = pv.data(); ws.pu = pu.data(); ws.pair_cap = pair_cap;
        ws.reg_cap = reg_cap;
        const int sw_cap = max_q + 32, sw_b = 1 << 16;
        std::vector<int32_t> swbuf(4 * (size_t)sw_cap);
        std::vector<uint64_t> swb(sw_b);
        ws.sw.H0 = swbuf.data(); ws.sw.H1 = swbuf.data() + sw_cap; ws.sw.E = swbuf.data() + 2 * sw_cap; ws.sw.Hmax = swbuf.data() + 3 * sw_cap;
        ws.sw.b = swb.data(); ws.sw.cap = sw_cap; ws.sw.cap_b = sw_b;
        std::vector<uint8_t> rev(max_q);
        ws.rev = rev.data();
        std::vector<AlnReg> wregs(2 * (size_t)(reg_cap + opt.max_matesw));
        // outputs
        out.reads.assign(n, ReadOut());
        size_t arena_cap = (size_t)n * 1024 + (1 << 20);
        size_t task_cap = (size_t)n * 2 + 1024;
        std::vector<AlnTask> tasks;
        for (;;) {
            out.arena.resize_uninit(arena_cap);
            memset(out.arena.data(), 0, arena_cap);
            tasks.assign(task_cap, AlnTask());
            unsigned long long used = 8; // offset 0 is reserved as "null"
            unsigned int n_tasks = 0;
            B.out = out.reads.data();
            B.arena.base = out.arena.data(); B.arena.used = &used; B.arena.cap = arena_cap;
            B.tasks.a = tasks.data(); B.tasks.n = &n_tasks; B.tasks.cap = (unsigned int)task_cap;
            if (pe && getenv("HOSTSIM_RESCUE_JOBS") && (long)max_len * opt.a < 250) {   // (the device takes this path for the 8-bit kernel only)
                // the rescue path of the device: pairs that need a rescue Smith-Waterman are set aside, all their jobs are
                // enumerated, computed by the lane machine (bsb_rescue.h) and the pairs are finalised over the results
                const int chunk = std::max(8, atoi(getenv("HOSTSIM_RESCUE_JOBS")) & ~7);
                std::vector<int> heavy;
                for (int p = 0; p < n / 2; ++p) if (stage_final_pe(opt, ix_, B, p, ws, wregs.data(), true)) heavy.push_back(p);
                std::vector<RescueJob> jobs(heavy.size() * 8 * (size_t)opt.max_matesw + 8);
                unsigned int n_jobs = 0;
                for (size_t h = 0; h < heavy.size(); ++h) {       // (the device counts first, scans, then writes each pair's block)
                    unsigned int cnt = 0;
                    RescueSink sink = {jobs.data() + n_jobs, &cnt, (unsigned int)(jobs.size() - n_jobs), (int32_t)h, 0};
                    stage_final_pe(opt, ix_, B, heavy[h], ws, wregs.data(), false, &sink, nullptr);
                    n_jobs += cnt;
                }
                if (n_jobs > jobs.size()) throw std::runtime_error("hostsim: rescue job list too small");
                std::vector<SwResult> res(n_jobs);
                const int cap_cells = ((max_len + 15) / 16) * 16;
                std::vector<uint32_t> tile(cap_cells + cap_cells / 8 + 8);
                std::vector<uint64_t> blist(256);
                for (unsigned int k = 0; k < n_jobs; ++k) {
                    SwLane<PackedRow<1>> L;
                    L.W.p = tile.data(); L.cap_cells = cap_cells; L.b = blist.data(); L.cap_b = (int)blist.size();
                    const RescueJob &jb = jobs[k];
                    L.begin(opt, ix_, jb, B.seq + B.seq_off[jb.mate_read]);
                    for (;;) {
                        if (L.state == SwLane<PackedRow<1>>::INIT) L.init_step(chunk);
                        else if (L.state == SwLane<PackedRow<1>>::ROWS) L.step(opt, chunk);
                        else if (L.state == SwLane<PackedRow<1>>::PASS_END) { if (L.end_pass()) break; }
                        else break;
                    }
                    if (L.err) throw std::runtime_error("hostsim: rescue Smith-Waterman failed");
                    res[k] = L.out;
                }
                size_t k0 = 0;
                for (size_t h = 0; h < heavy.size(); ++h) {       // jobs of a pair are contiguous here (serial enumeration)
                    size_t k1 = k0;
                    while (k1 < n_jobs && jobs[k1].pair == (int32_t)h) ++k1;
                    const RescuePre pre = {jobs.data() + k0, res.data() + k0, (int)(k1 - k0)};
                    stage_final_pe(opt, ix_, B, heavy[h], ws, wregs.data(), false, nullptr, &pre);
                    k0 = k1;
                }
            } else
            if (pe) for (int p = 0; p < n / 2; ++p) stage_final_pe(opt, ix_, B, p, ws, wregs.data());
            else for (int r = 0; r < n; ++r) stage_final_se(opt, ix_, B, r, ws, wregs.data());
            if (n_tasks <= task_cap) for (unsigned int k = 0; k < n_tasks; ++k) stage_task(opt, ix_, B, k, ws);
            bool ovf = n_tasks > task_cap || used > arena_cap;
            for (int r = 0; r < n; ++r) if (out.reads[r].err == ERR_ARENA_OVERFLOW) ovf = true;
            if (!ovf) break;
            if (n_tasks > task_cap) task_cap = (size_t)n_tasks + 1024;
            arena_cap *= 2;
        }
        for (int r = 0; r < n; ++r)
            if (out.reads[r].err) throw std::runtime_error("hostsim: read " + b.name(r) + " failed with error code " + std::to_string(out.reads[r].err));
        out.have_bam = false;
        if (out.want_bam && !idx_.any_alt && b.comments.empty() && n) {
            // the device's BAM stage, emulated (bam_emulation.h): SAM text of every entry as the device formatter leaves it, then the
            // arbiter, the records and the BGZF blocks through the functions the kernels run
            MemArgs ma;
            ma.opt = opt; ma.rg_id = out.rg_id;
            std::string text, one;
            std::vector<uint32_t> text_off(n + 1, 0);
            std::vector<SamStats> stats(n);
            for (int i = 0; i < n; ++i) {
                EntryStats st;
                format_entry(ma, idx_, b, i, out, one, st);
                text += one;
                text_off[i + 1] = (uint32_t)text.size();
                stats[i].alignment_score = st.alignment_score; stats[i].mapped = st.mapped; stats[i].bs_conflict = st.bs_conflict; stats[i].crick = st.crick; stats[i].paired = st.paired;
            }
            BamBatchIn in;
            in.names = b.names.data(); in.name_off = b.name_off.data(); in.first = b.first.data(); in.read_group = b.read_group.data();
            in.bases = b.bases.data(); in.qual = b.qual.data(); in.seq_off = b.seq_off.data(); in.has_qual = b.has_qual.data();
            in.text = text.data(); in.text_off = text_off.data(); in.stats = stats.data(); in.n = n;
            in.ctg_text = idx_.ctg_text.data(); in.ctg_name_off = idx_.ctg_name_off.data(); in.ctg_sorted = idx_.ctg_sorted.data(); in.n_ctg = (int)idx_.ctg_sorted.size();
            std::vector<uint8_t> bgzf;
            MapCounters ctr;
            bam_batch_emulated(in, bgzf, ctr, out.bam_raw_bytes, out.bam_records);
            out.bam.resize_uninit(bgzf.size());
            if (!bgzf.empty()) memcpy(out.bam.data(), bgzf.data(), bgzf.size());
            const unsigned long long *f = reinterpret_cast<const unsigned long long *>(&ctr);
            for (int k = 0; k < 8; ++k) out.bam_counts[k] = f[k];
            out.have_bam = true;
        }
    }
private:
    const HostIndex &idx_;
    IndexView ix_;
    std::vector<double> log_tab_;
    std::vector<uint32_t> occ32_;
};

// HOSTSIM_DEVICES=n: the pipeline's multi-device path (batch b -> device b mod n, one slot per device, results collected in
// input order) over n CPU "devices" -- the host-side logic of bsb_mem_main_multi without a GPU
class FakeDevices : public BatchAligner {
public:
    FakeDevices(BatchAligner &one, int n) : one_(one), n_(n), used_(n, 0) {}
    int devices() const override { return n_; }
    int slots() const override { return n_; }
    void align(const Opt &opt, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int slot = 0) override
    {
        { std::lock_guard<std::mutex> l(m_); ++used_[slot]; }
        one_.align(opt, b, n_processed, pes0, out, 0);
    }
    void report() { for (int d = 0; d < n_; ++d) fprintf(stderr, "[D::hostsim] device %d aligned %ld batches\n", d, used_[d]); }
private:
    BatchAligner &one_; int n_; std::vector<long> used_; std::mutex m_;
};

// hostsim readbench <chunk bases> <parse threads> <fill threads> <undirectional 0|1> <in1.fq> [in2.fq]: the reader alone (cut + batch + copy)
static int readbench(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: hostsim readbench <chunk> <parse threads> <fill threads> <undirectional> <in1.fq> [in2.fq]\n"); return 1; }
    const int64_t chunk = atoll(argv[2]);
    const int pt = atoi(argv[3]), ft = atoi(argv[4]), un = atoi(argv[5]);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    FastxReader r1(argv[6], pt, un != 0);
    std::unique_ptr<FastxReader> r2;
    if (argc > 7) r2.reset(new FastxReader(argv[7], pt, un != 0));
    BatchPlan plan;
    ReadBatch batch;
    long n = 0, nb = 0;
    double t_plan = 0, t_fill = 0;
    auto now = [] { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
    for (;;) {
        double t0 = now();
        if (!plan_batch(chunk, &r1, r2.get(), un, 0.1f, plan)) break;
        double t1 = now();
        fill_batch(plan, &r1, r2.get(), false, ft, batch);
        t_plan += t1 - t0; t_fill += now() - t1;
        n += batch.n; ++nb;
    }
    clock_gettime(CLOCK_MONOTONIC, &b);
    const double s = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    printf("%ld entries in %ld batches, %.3f s: %.1f M entries/s (plan %.3f s, fill %.3f s, serialised here)\n", n, nb, s, n / s / 1e6, t_plan, t_fill);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && strcmp(argv[1], "readbench") == 0) return readbench(argc, argv);
    if (argc < 2 || strcmp(argv[1], "mem") != 0) { fprintf(stderr, "usage: hostsim mem [options] <idxbase> <in1.fq> [in2.fq]\n"); return 1; }
    try {
        MemArgs ma;
        std::string err;
        if (parse_mem_args(argc - 1, argv + 1, ma, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        ma.pg_line = "@PG\tID:bwa\tPN:bwa\tVN:hostsim";
        HostIndex idx;
        idx.load(ma.idxbase);
        if (ma.ignore_alt) for (auto &a : idx.anns) a.is_alt = 0;
        HostSimAligner al(idx);
        RunSummary sum;
        if (const char *e = getenv("HOSTSIM_BAM")) {   // HOSTSIM_BAM=<path>: the run's BAM file; HOSTSIM_BAM_HOST=1: through the host encoder (zlib)
            BamWriter bw(e, 2, -1);
            bw.accept_device_blocks(!getenv("HOSTSIM_BAM_HOST"));
            const int rc = run_mem(ma, idx, al, nullptr, stderr, &sum, &bw);
            bw.close();
            return rc;
        }
        if (const char *e = getenv("HOSTSIM_DEVICES")) {
            FakeDevices multi(al, std::max(1, atoi(e)));
            const int rc = run_mem(ma, idx, multi, stdout, stderr, &sum);
            multi.report();
            return rc;
        }
        return run_mem(ma, idx, al, stdout, stderr, &sum);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
