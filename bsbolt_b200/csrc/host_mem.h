// host_mem.h -- host side of `bwa mem` as BSBolt runs it: option parsing, SAM header, SAM text
// assembly, the bisulfite read-group arbiter and the batch loop. The GPU work sits behind the
// BatchAligner interface (the seam is the reference's mem_process_seqs, bwamem.c:1319).
//   parse_mem_args()  <- main_mem option loop + update_a       (fastmap.c:78-317)
//   sam_header()      <- bwa_print_sam_hdr                      (bwa.c:530-553)
//   format_entry()    <- mem_aln2sam                            (bwamem.c:829-1053)
//   SamSorter         <- samSorter + wrapper                    (bs_sorter.cpp, bs_sorter_wrapper.cpp)
//   estimate_pestat() <- mem_pestat from per-pair candidates    (bwamem_pair.c:46-109)
//   run_mem()         <- main_mem + process() pipeline          (fastmap.c:10-76, 319-363)
#pragma once
#include "bsb_sam.h"
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "bsb_types.h"
#include "host_index.h"
#include "host_reads.h"

namespace bsb {

struct MemArgs {
    Opt opt;
    bool copy_comment = false, ignore_alt = false, no_mt_io = false;
    int fixed_chunk_size = -1;
    int verbose = 3;
    std::string hdr_line; bool have_hdr = false;
    std::string rg_id;
    bool have_pes0 = false;
    PeStat pes0[4];
    std::string idxbase, fq1, fq2, out_path;
    std::string pg_line;         // "@PG\tID:bwa\tPN:bwa\tVN:...\tCL:..."
    // multi-GPU: this process aligns batches b with b % shard_count == shard_index; every shard
    // scans the whole input so that batch boundaries and n_processed are those of a single run
    int shard_index = 0, shard_count = 1;
    std::string shard_parts;     // sidecar listing (batch id, byte offset, length) of this shard's SAM output
    int64_t actual_chunk_size() const { return fixed_chunk_size > 0 ? fixed_chunk_size : (int64_t)opt.chunk_size * opt.n_threads; }
};

void opt_init(Opt &o);                              // mem_opt_init (bwamem.c:50-88)
void fill_scmat(int a, int b, int8_t mat[25]);      // bwa_fill_scmat (bwa.c:169-178)
// argv as given to `bwa mem` (argv[0] == "mem"). Returns 0 on success, 1 on usage error (message in err).
int parse_mem_args(int argc, char **argv, MemArgs &ma, std::string &err);

struct BatchResult {
    PinArray<ReadOut> reads;
    RawBuf arena;
    PeStat pes[4];
    // hot-path accounting for the benchmark (CUDA-event timings in ms on the aligner's stream)
    // ms_stage: 0 H2D, 1 convert, 2 seed, 3 scan+SA lookup, 4 chain, 5 extend, 6 pair stats, 7 finalise
    double ms_h2d = 0, ms_kernels = 0, ms_d2h = 0, ms_stage[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t n_seeds = 0, h2d_bytes = 0, d2h_bytes = 0;
    double ms_select = 0, ms_tasks = 0; uint64_t n_tasks = 0;   // split of stage 7: record selection / alignment tasks
    // counted work of the batch: FM-index extensions of the seeding kernel, how many of them read two occ blocks, bytes per
    // occ block of the layout in use; DP cells filled by the extension kernel (0 where the warp-per-read form ran)
    uint64_t n_fm_ext = 0, n_fm_two_block = 0, n_fm_two_block_ref = 0, n_ext_cells = 0; int fm_block_bytes = 0;
    uint64_t n_rescue_pairs = 0, n_rescue_jobs = 0;   // pairs that needed mate-rescue Smith-Waterman; Smith-Waterman jobs computed for them
    // SAM text formatted on the device (bsb_sam.h): requested by the caller with want_text (+ the read-group id, if any);
    // when have_text comes back true, `text` holds the records of all entries back to back (entry i = bytes
    // [text_off[i], text_off[i+1])), `stats` the per-entry statistics the arbiter needs, and `arena` was not copied back.
    bool want_text = false, have_text = false;
    std::string rg_id;
    RawBuf text; PinArray<uint32_t> text_off; PinArray<SamStats> stats;
    double ms_text = 0;
    // BAM on the device (bsb_bam.h, bsb_deflate.h): requested with want_bam (the caller writes a BAM file and no comments are
    // appended). When have_bam comes back true the arbiter ran on the device, `bam` holds the finished BGZF blocks of the batch's
    // records in input order -- the host only appends them to the file -- and neither text nor arena were copied back.
    bool want_bam = false, have_bam = false;
    RawBuf bam;
    uint64_t bam_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // reads, alignments, W_C2T, W_G2A, C_C2T, C_G2A, unaligned, BS-ambiguous (the BSStat lines)
    uint64_t bam_raw_bytes = 0, bam_records = 0, bam_blocks = 0;
    double ms_bam = 0;
    std::string log_text;        // log lines of this batch (mem_pestat), printed by the pipeline as one block
};

class BatchAligner {
public:
    virtual ~BatchAligner() {}
    // Aligns one batch. n_processed is the number of bseq entries in all earlier batches.
    // Throws std::runtime_error on device errors or scratch overflow (never truncates silently).
    // `slot` < slots() selects one of the aligner's independent batch contexts; calls on different slots may run
    // concurrently from different host threads.
    virtual void align(const Opt &opt, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int slot = 0) = 0;
    virtual int slots() const { return 1; }
    // An aligner may drive several GPUs (index replicated, reads sharded by batch, no collective): slots
    // [d * slots() / devices(), (d + 1) * slots() / devices()) belong to device d. The pipeline hands batch b to device
    // b mod devices() and collects the results in input order.
    virtual int devices() const { return 1; }
    // Measurement aid: uploads the inputs of `b` ahead of time (b.dev_input) to the device that owns `slot`, so that a
    // following align() on a slot of that device starts with its batch already resident in HBM; unload() frees them.
    virtual void preload(ReadBatch &, int /*slot*/ = 0) {}
    virtual void unload(ReadBatch &) {}
};

struct EntryStats { int alignment_score = 0, mapped = 0, bs_conflict = 0, crick = 0, paired = 0; };

struct SamView;
SamView make_sam_view(const MemArgs &ma, const HostIndex &idx, const ReadBatch &b, const BatchResult &res);
std::string sam_header(const HostIndex &idx, const MemArgs &ma);
void format_entry(const MemArgs &ma, const HostIndex &idx, const ReadBatch &b, int i, const BatchResult &res,
                  std::string &sam, EntryStats &st);

struct MapStats {
    long reads = 0, alignments = 0, wc2t = 0, wg2a = 0, cc2t = 0, cg2a = 0, unaligned = 0, bs_ambiguous = 0;
    void add(const MapStats &o);
};

// Arbitrates between the two conversion-pattern groups of each read name and appends the chosen
// SAM text to `out`, in input order. One instance per batch, like the reference.
void sam_sort_batch(const ReadBatch &b, std::vector<std::string> &sam, std::vector<EntryStats> &st,
                    std::string &out, MapStats &stats);

// Insert-size distribution from the per-pair candidates (dir in [0,4) or -1, insert size).
// The [M::mem_pestat] lines are appended to *log_text when given (the caller prints them under its log lock), else written to stderr.
void estimate_pestat(const Opt &opt, const std::vector<int8_t> &dir, const std::vector<int64_t> &isize, PeStat pes[4], int verbose, std::string *log_text = nullptr);

// glibc-evaluated tables shipped to the device (see MathTab in bsb_final.h)
void build_log_table(std::vector<double> &t, int n);
void build_pair_table(const Opt &opt, const PeStat pes[4], std::vector<double> &t, int off[4]);

struct RunSummary {
    MapStats stats; long n_batches = 0; long n_entries = 0; double sec_total = 0, sec_align = 0;
    double sec_read = 0, sec_format = 0, sec_write = 0; // busy time of the reader / formatter / output stages
    double sec_plan = 0, sec_fill = 0;                  // the reader's two overlapped halves: cutting batches (incl. waiting for the parsers) / copying them
    double sec_resident = 0;  // BSB_RESIDENT_BENCH: wall time from "all batches resident on the device" to "last batch aligned"
    double ms_h2d = 0, ms_kernels = 0, ms_d2h = 0, ms_stage[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t n_seeds = 0, h2d_bytes = 0, d2h_bytes = 0, n_tasks = 0;
    uint64_t n_fm_ext = 0, n_fm_two_block = 0, n_fm_two_block_ref = 0, n_ext_cells = 0; int fm_block_bytes = 0;
    uint64_t n_rescue_pairs = 0, n_rescue_jobs = 0;
    double ms_select = 0, ms_tasks = 0;
    double ms_text = 0, ms_bam = 0; uint64_t bam_raw_bytes = 0, bam_bgzf_bytes = 0, bam_blocks = 0;
    void add_timing(const BatchResult &r)
    {
        ms_text += r.ms_text; ms_bam += r.ms_bam;
        if (r.have_bam) { bam_raw_bytes += r.bam_raw_bytes; bam_bgzf_bytes += r.bam.size(); bam_blocks += r.bam_blocks; }
        ms_select += r.ms_select; ms_tasks += r.ms_tasks; n_tasks += r.n_tasks;
        ms_h2d += r.ms_h2d; ms_kernels += r.ms_kernels; ms_d2h += r.ms_d2h;
        for (int k = 0; k < 8; ++k) ms_stage[k] += r.ms_stage[k];
        n_seeds += r.n_seeds; h2d_bytes += r.h2d_bytes; d2h_bytes += r.d2h_bytes;
        n_fm_ext += r.n_fm_ext; n_fm_two_block += r.n_fm_two_block; n_fm_two_block_ref += r.n_fm_two_block_ref; n_ext_cells += r.n_ext_cells; if (r.fm_block_bytes) fm_block_bytes = r.fm_block_bytes;
        n_rescue_pairs += r.n_rescue_pairs; n_rescue_jobs += r.n_rescue_jobs;
    }
};

// The whole `bwa mem` run. SAM goes to `out`, log/BSStat lines to `log`.
class BamWriter;   // host_bam.h
// SAM text goes to `out`, or -- when `bam` is given -- BAM records go to the BamWriter and `out` may be NULL
int run_mem(const MemArgs &ma, const HostIndex &idx, BatchAligner &aligner, FILE *out, FILE *log, RunSummary *summary, BamWriter *bam = nullptr);

} // namespace bsb
