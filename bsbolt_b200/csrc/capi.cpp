// capi.cpp -- extern "C" surface declared in include/bsbolt_b200.h
#include "../../include/bsbolt_b200.h"
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime_api.h>
#include "bsb_cuda.h"
#include "bsb_index_build.h"
#include "host_bam.h"
#include <thread>

using namespace bsb;

static thread_local std::string g_err;
static thread_local std::string g_hdr;

struct bsb_index {
    std::shared_ptr<HostIndex> host;   // shared by the copies of one index on several devices (bsb_index_clone)
    std::unique_ptr<CudaAligner> aligner;
    int device = 0;
};

struct bsb_batch {
    bsb_index *idx = nullptr;
    MemArgs ma;
    ReadBatch reads;
    BatchResult res;
    bool aligned = false;
    std::string sam;
};

static void fill_stats(bsb_run_stats_t *s, const RunSummary &sum, const CudaAligner *al, long launches = -1)
{
    if (!s) return;
    s->total_reads = sum.stats.reads; s->total_alignments = sum.stats.alignments;
    s->w_c2t = sum.stats.wc2t; s->w_g2a = sum.stats.wg2a; s->c_c2t = sum.stats.cc2t; s->c_g2a = sum.stats.cg2a;
    s->unaligned = sum.stats.unaligned; s->bs_ambiguous = sum.stats.bs_ambiguous;
    s->n_batches = sum.n_batches; s->n_entries = sum.n_entries; s->sec_total = sum.sec_total; s->sec_align = sum.sec_align;
    s->ms_h2d = sum.ms_h2d; s->ms_kernels = sum.ms_kernels; s->ms_d2h = sum.ms_d2h;
    for (int k = 0; k < 8; ++k) s->ms_stage[k] = sum.ms_stage[k];
    s->n_seeds = (int64_t)sum.n_seeds; s->h2d_bytes = (int64_t)sum.h2d_bytes; s->d2h_bytes = (int64_t)sum.d2h_bytes;
    s->kernel_launches = launches >= 0 ? launches : al ? al->kernel_launches() : 0;
    s->sec_read = sum.sec_read; s->sec_format = sum.sec_format; s->sec_write = sum.sec_write;
    s->ms_select = sum.ms_select; s->ms_tasks = sum.ms_tasks; s->n_tasks = (int64_t)sum.n_tasks;
    s->sec_resident = sum.sec_resident;
    s->fm_extensions = (int64_t)sum.n_fm_ext; s->fm_two_block = (int64_t)sum.n_fm_two_block; s->fm_block_bytes = sum.fm_block_bytes;
    s->dp_cells_extend = (int64_t)sum.n_ext_cells; s->fm_two_block_ref = (int64_t)sum.n_fm_two_block_ref;
    s->sec_plan = sum.sec_plan; s->sec_fill = sum.sec_fill; s->rescue_pairs = (int64_t)sum.n_rescue_pairs; s->rescue_jobs = (int64_t)sum.n_rescue_jobs;
    s->ms_text = sum.ms_text; s->ms_bam = sum.ms_bam; s->bam_raw_bytes = (int64_t)sum.bam_raw_bytes; s->bam_bgzf_bytes = (int64_t)sum.bam_bgzf_bytes; s->bam_blocks = (int64_t)sum.bam_blocks;
}

extern "C" {

const char *bsb_version(void) { return "bsbolt_b200 0.1 (BSB-1.2.1-BWA-fork-0.7.17 semantics, sm_100a)"; }
const char *bsb_last_error(void) { return g_err.c_str(); }

size_t bsb_run_stats_size(void) { return sizeof(bsb_run_stats_t); }

int bsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

bsb_index_t *bsb_index_load(const char *idxbase, int device)
{
    try {
        std::unique_ptr<bsb_index> ix(new bsb_index);
        ix->device = device;
        ix->host.reset(new HostIndex);
        ix->host->load(idxbase);
        ix->aligner.reset(new CudaAligner(*ix->host, device));
        return ix.release();
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

bsb_index_t *bsb_index_clone(const bsb_index_t *src, int device)
{
    try {
        if (!src) throw std::runtime_error("[E::bsb_index_clone] index is NULL");
        std::unique_ptr<bsb_index> ix(new bsb_index);
        ix->device = device;
        ix->host = src->host;
        ix->aligner.reset(new CudaAligner(*ix->host, device));
        return ix.release();
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void bsb_index_free(bsb_index_t *idx) { delete idx; }
int64_t bsb_index_hbm_bytes(const bsb_index_t *idx) { return idx ? (int64_t)idx->aligner->index_bytes() : 0; }
int bsb_index_n_contigs(const bsb_index_t *idx) { return idx ? (int)idx->host->contigs.size() : 0; }

static std::string make_pg(int argc, char **argv)
{
    std::string pg = "@PG\tID:bwa\tPN:bwa\tVN:BSB-1.2.1-BWA-fork-0.7.17-b200\tCL:bsbolt_b200";
    for (int i = 0; i < argc; ++i) { pg += ' '; pg += argv[i]; }
    return pg;
}

static int host_thread_share()
{
    const int n = host_core_share();
    return std::max(1, std::min(n - 2, 64));   // the device threads sleep on events; leave two cores to the FASTQ reader
}

static int mem_main_impl(bsb_index_t *idx, int device, int argc, char **argv, int out_fd, const char *bam_path, int bam_threads, int bam_level,
                         int log_fd, bsb_run_stats_t *stats, bsb_index_t *const *more = nullptr, int n_more = 0)
{
    FILE *out = nullptr, *log = nullptr;
    bsb_index_t *own = nullptr;
    std::unique_ptr<BamWriter> bam;
    int ret = 1;
    try {
        MemArgs ma;
        std::string err;
        if (parse_mem_args(argc, argv, ma, err)) throw std::runtime_error(err);
        ma.pg_line = make_pg(argc, argv);
        if (!idx) {
            own = bsb_index_load(ma.idxbase.c_str(), device);
            if (!own) throw std::runtime_error(g_err);
            idx = own;
        }
        int lfd = dup(log_fd);
        log = fdopen(lfd, "w");
        if (log) setvbuf(log, nullptr, _IOLBF, 1 << 16);   // whole lines: the descriptor may be shared with other writers (stderr)
        if (bam_path) bam.reset(new BamWriter(bam_path, bam_threads > 0 ? bam_threads : host_thread_share(), bam_level));
        if (bam) bam->accept_device_blocks(bam_level < 0 && !getenv("BSB_BAM_HOST"));   // default level: compressed on the device (bsb_deflate.h)
        else {
            int ofd = dup(out_fd);
            if (!ma.out_path.empty()) { out = fopen(ma.out_path.c_str(), "wb"); if (ofd >= 0) close(ofd); }
            else out = fdopen(ofd, "w");
            if (out) setvbuf(out, nullptr, _IOFBF, 1 << 22);
        }
        if ((!out && !bam) || !log) throw std::runtime_error("[E::bsb_mem_main] cannot open the output streams");
        // -j (fastmap.c: bns->anns[i].is_alt = 0 for every contig): a no-op unless the database carries an .alt file
        if (ma.ignore_alt && idx->host->any_alt)
            throw std::runtime_error("[E::bsb_mem_main] -j on a database with ALT contigs needs an index loaded without the ALT marks; not supported with a resident index");
        idx->aligner->verbose = ma.verbose;
        RunSummary sum;
        if (n_more > 0) {   // several devices: one reader, batches dealt round-robin, output in input order
            std::vector<CudaAligner *> per_device(1, idx->aligner.get());
            for (int k = 0; k < n_more; ++k) {
                if (!more[k] || more[k]->host.get() != idx->host.get())
                    throw std::runtime_error("[E::bsb_mem_main_multi] every index must be a bsb_index_clone of the first");
                per_device.push_back(more[k]->aligner.get());
            }
            MultiAligner multi(per_device);
            multi.set_verbose(ma.verbose);
            const long launches0 = multi.kernel_launches();
            ret = run_mem(ma, *idx->host, multi, out, log, &sum, bam.get());
            if (bam) { bam->close(); sum.sec_write += bam->sec_busy(); }
            fill_stats(stats, sum, nullptr, multi.kernel_launches() - launches0);
        } else {
        ret = run_mem(ma, *idx->host, *idx->aligner, out, log, &sum, bam.get());
        if (bam) { bam->close(); sum.sec_write += bam->sec_busy(); }
        fill_stats(stats, sum, idx->aligner.get());
        }
    } catch (const std::exception &e) {
        g_err = e.what();
        if (log) fprintf(log, "%s\n", e.what());
        ret = 1;
    }
    bam.reset();
    if (out) fclose(out);
    if (log) fclose(log);
    if (own) bsb_index_free(own);
    return ret;
}

int bsb_mem_main(bsb_index_t *idx, int device, int argc, char **argv, int out_fd, int log_fd, bsb_run_stats_t *stats)
{
    return mem_main_impl(idx, device, argc, argv, out_fd, nullptr, 0, -1, log_fd, stats);
}

int bsb_mem_main_bam(bsb_index_t *idx, int device, int argc, char **argv, const char *bam_path, int threads, int level, int log_fd,
                     bsb_run_stats_t *stats)
{
    if (!bam_path) { g_err = "[E::bsb_mem_main_bam] bam_path is NULL"; return 1; }
    return mem_main_impl(idx, device, argc, argv, -1, bam_path, threads, level, log_fd, stats);
}

int bsb_mem_main_multi(bsb_index_t *const *idx, int n_idx, int argc, char **argv, int out_fd, int log_fd, bsb_run_stats_t *stats)
{
    if (!idx || n_idx < 1 || !idx[0]) { g_err = "[E::bsb_mem_main_multi] no index"; return 1; }
    return mem_main_impl(idx[0], idx[0]->device, argc, argv, out_fd, nullptr, 0, -1, log_fd, stats, idx + 1, n_idx - 1);
}

int bsb_mem_main_multi_bam(bsb_index_t *const *idx, int n_idx, int argc, char **argv, const char *bam_path, int threads, int level, int log_fd,
                           bsb_run_stats_t *stats)
{
    if (!idx || n_idx < 1 || !idx[0]) { g_err = "[E::bsb_mem_main_multi_bam] no index"; return 1; }
    if (!bam_path) { g_err = "[E::bsb_mem_main_multi_bam] bam_path is NULL"; return 1; }
    return mem_main_impl(idx[0], idx[0]->device, argc, argv, -1, bam_path, threads, level, log_fd, stats, idx + 1, n_idx - 1);
}

int64_t bsb_stream_bam(int in_fd, const char *bam_path, int threads, int level)
{
    try {
        if (!bam_path) throw std::runtime_error("[E::bsb_stream_bam] bam_path is NULL");
        return (int64_t)stream_bam(in_fd, bam_path, threads > 0 ? threads : host_thread_share(), level);
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

bsb_batch_t *bsb_batch_create(bsb_index_t *idx, int opt_argc, char **opt_argv, int n, const bsb_read_t *r1, const bsb_read_t *r2)
{
    try {
        if (!idx) throw std::runtime_error("[E::bsb_batch_create] index is NULL");
        std::unique_ptr<bsb_batch> b(new bsb_batch);
        b->idx = idx;
        // parse_mem_args wants the positional arguments; supply placeholders
        std::vector<char *> av(opt_argv, opt_argv + opt_argc);
        char p0[] = "idx", p1[] = "r1", p2[] = "r2";
        av.push_back(p0); av.push_back(p1);
        if (r2) av.push_back(p2);
        std::string err;
        if (parse_mem_args((int)av.size(), av.data(), b->ma, err)) throw std::runtime_error(err);
        b->ma.verbose = 0;
        b->reads.clear();
        auto trim = [](std::string &s) {
            size_t l = s.size();
            if (l > 2 && s[l - 2] == '/' && s[l - 1] >= '0' && s[l - 1] <= '9') s.resize(l - 2);
        };
        const Opt &opt = b->ma.opt;
        for (int i = 0; i < n; ++i) { // same entry construction as read_batch() (bwa.c:73-145)
            FastxRecord k1, k2;
            k1.name = r1[i].name; k1.seq = r1[i].seq; if (r1[i].comment) k1.comment = r1[i].comment; if (r1[i].qual) k1.qual = r1[i].qual;
            trim(k1.name);
            if (r2) { k2.name = r2[i].name; k2.seq = r2[i].seq; if (r2[i].comment) k2.comment = r2[i].comment; if (r2[i].qual) k2.qual = r2[i].qual; trim(k2.name); }
            int pattern = 0, both = 0;
            if (opt.undirectional) {
                int t = r2 ? assess_conversion(k1.seq.data(), k1.seq.size(), k2.seq.data(), k2.seq.size(), 1, opt.substitution_proportion)
                           : assess_conversion(k1.seq.data(), k1.seq.size(), k1.seq.data(), k1.seq.size(), 0, opt.substitution_proportion);
                if (t == 2) both = 1; else pattern = t;
            }
            b->reads.add(k1, b->ma.copy_comment, 0, 0, pattern);
            if (r2) b->reads.add(k2, b->ma.copy_comment, 1, 0, pattern ? 0 : 1);
            if (both) {
                b->reads.add(k1, b->ma.copy_comment, 0, 1, 1);
                if (r2) b->reads.add(k2, b->ma.copy_comment, 1, 1, 0);
            }
        }
        return b.release();
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

int bsb_batch_align(bsb_batch_t *b, int64_t n_processed, bsb_run_stats_t *stats)
{
    try {
        if (!b) throw std::runtime_error("[E::bsb_batch_align] batch is NULL");
        b->idx->aligner->verbose = 0;
        b->idx->aligner->align(b->ma.opt, b->reads, n_processed, b->ma.have_pes0 ? b->ma.pes0 : nullptr, b->res);
        if (!b->res.log_text.empty()) { fputs(b->res.log_text.c_str(), stderr); b->res.log_text.clear(); }   // [M::mem_pestat] lines, like the reference
        b->aligned = true;
        if (stats) {
            RunSummary sum;
            sum.n_batches = 1; sum.n_entries = b->reads.n;
            sum.add_timing(b->res);
            fill_stats(stats, sum, b->idx->aligner.get());
        }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

int bsb_batch_sam(bsb_batch_t *b, const char **sam, size_t *len, bsb_run_stats_t *stats)
{
    try {
        if (!b || !b->aligned) throw std::runtime_error("[E::bsb_batch_sam] batch has not been aligned");
        std::vector<std::string> lines(b->reads.n);
        std::vector<EntryStats> st(b->reads.n);
        for (int i = 0; i < b->reads.n; ++i) format_entry(b->ma, *b->idx->host, b->reads, i, b->res, lines[i], st[i]);
        b->sam.clear();
        MapStats ms;
        sam_sort_batch(b->reads, lines, st, b->sam, ms);
        if (sam) *sam = b->sam.c_str();
        if (len) *len = b->sam.size();
        if (stats) {
            RunSummary sum;
            sum.stats = ms; sum.n_batches = 1; sum.n_entries = b->reads.n;
            sum.add_timing(b->res);
            fill_stats(stats, sum, b->idx->aligner.get());
        }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

int bsb_batch_n_entries(const bsb_batch_t *b) { return b ? b->reads.n : 0; }
void bsb_batch_free(bsb_batch_t *b) { delete b; }

int bsb_random_sector_peak(int device, size_t footprint_bytes, double *gbs_independent, double *gbs_chase)
{
    try {
        double a = 0, b = 0;
        random_sector_peak(device, footprint_bytes, &a, &b);
        if (gbs_independent) *gbs_independent = a;
        if (gbs_chase) *gbs_chase = b;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

const char *bsb_sam_header(bsb_index_t *idx, int argc, char **argv)
{
    try {
        if (!idx) throw std::runtime_error("[E::bsb_sam_header] index is NULL");
        MemArgs ma;
        std::string err;
        if (parse_mem_args(argc, argv, ma, err)) throw std::runtime_error(err);
        ma.pg_line = make_pg(argc, argv);
        g_hdr = sam_header(*idx->host, ma);
        return g_hdr.c_str();
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

int bsb_index_build(const char *fasta, const char *prefix, int device, double *device_ms)
{
    try {
        IndexBuildStats st;
        index_build(fasta, prefix, device, &st);
        if (device_ms) *device_ms = st.ms_device;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

} // extern "C"
