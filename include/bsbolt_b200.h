/* bsbolt_b200.h -- C ABI of the B200-native bisulfite aligner (libbsbolt_b200.so).
 *
 * Drop-in boundary for the `bsbolt Align` hot path. The reference reaches its aligner through a
 * process boundary (`bwa mem` argv + pipes, bsbolt/Align/AlignReads.py:43-87); in-process the seam
 * is main_mem / mem_process_seqs. Each entry point below names the reference interface it replaces
 * (paths relative to bsbolt/External/BWA/ of NuttyLogic/BSBolt v1.6.0).
 *
 * Plain C types only: pointers, sizes, ints. All functions are safe to call from any one thread at
 * a time per bsb_index_t. On failure functions return NULL / non-zero and bsb_last_error() holds
 * the message. There is no CPU fallback: without a CUDA device bsb_index_load fails.
 */
#ifndef BSBOLT_B200_H
#define BSBOLT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsb_index bsb_index_t; /* index files of `bsbolt Index` resident in one GPU's HBM */
typedef struct bsb_batch bsb_batch_t; /* one batch of reads (reference: bseq1_t[] of one kt_pipeline step) */

typedef struct {
    /* BSStat counters, summed over batches (bs_sorter_wrapper.cpp:13-23) */
    int64_t total_reads, total_alignments, w_c2t, w_g2a, c_c2t, c_g2a, unaligned, bs_ambiguous;
    int64_t n_batches, n_entries;   /* batches processed; bseq entries (reads x conversion patterns) */
    double sec_total, sec_align;    /* host wall clock: whole run; inside the batch aligner */
    /* CUDA-event sums over batches, milliseconds, on the aligner's stream:
       stage 0 H2D, 1 convert, 2 SMEM seeding, 3 scan + SA lookup, 4 chaining, 5 extension,
       6 pair statistics (incl. host round trip), 7 finalisation (CIGAR/MD/XB/pairing) */
    double ms_h2d, ms_kernels, ms_d2h, ms_stage[8];
    int64_t n_seeds, h2d_bytes, d2h_bytes, kernel_launches;
    double sec_read, sec_format, sec_write; /* busy time of the host stages: FASTQ batching, SAM text, output */
    double ms_select, ms_tasks;             /* stage 7 split: record selection/pairing kernel, alignment-task kernel */
    int64_t n_tasks;                        /* global alignments queued (records + XA entries) */
    double sec_resident;                    /* BSB_RESIDENT_BENCH=1 only (measurement): wall time from "every batch's input
                                               resident in HBM" to "last batch left the device"; 0 otherwise */
    /* algorithmic work counted on the device during the run (roofline numerators): FM-index extensions of the seeding kernel
       (bwt_extend calls of the reference, bwt.c:262-275), how many of them touch two occ blocks, bytes of one occ block in the
       layout used; cells of the banded extension DP (ksw_extend2 inner loop, ksw.c:439-454) */
    int64_t fm_extensions, fm_two_block, fm_block_bytes, dp_cells_extend;
    int64_t fm_two_block_ref;              /* ... of them that touch two of the reference's 64-byte occ blocks (bwt.h:72-78) */
    double sec_plan, sec_fill;             /* sec_read split: cutting the batches (incl. waiting for the parser threads) / copying them */
    int64_t rescue_pairs, rescue_jobs;     /* pairs that needed mate-rescue Smith-Waterman (mem_matesw) and the Smith-Waterman jobs run for them */
    /* output stages on the device, CUDA-event sums in ms: SAM text of the records; for a BAM file at the default level the
       arbiter + BAM records + BGZF deflate that follow it (bsb_mem_main_bam), with the uncompressed and compressed sizes */
    double ms_text, ms_bam;
    int64_t bam_raw_bytes, bam_bgzf_bytes, bam_blocks;
} bsb_run_stats_t;

typedef struct {
    const char *name;     /* NUL-terminated; "/1" "/2" suffixes are trimmed like bwa.c:28-32 */
    const char *comment;  /* may be NULL */
    const char *seq;      /* ASCII bases as in the FASTQ (unconverted) */
    const char *qual;     /* may be NULL */
} bsb_read_t;

const char *bsb_version(void);
const char *bsb_last_error(void);
int bsb_device_count(void);                       /* number of CUDA devices, 0 if none */
size_t bsb_run_stats_size(void);                  /* sizeof(bsb_run_stats_t) of the library: a binding checks its own layout against it */

/* replaces bwa_idx_load(hint, BWA_IDX_ALL) (bwa.c:407-443): reads <idxbase>.bwt .sa .ann .amb .pac .opac
 * and uploads them to HBM of `device` */
bsb_index_t *bsb_index_load(const char *idxbase, int device);
/* one more resident copy of a loaded index, on another device; the host-side tables are shared. The reference's
 * multi-threaded run shares one bwaidx_t between its worker threads (fastmap.c:319-352); on GPUs the index is replicated
 * per device and the batches are dealt out. */
bsb_index_t *bsb_index_clone(const bsb_index_t *src, int device);
void bsb_index_free(bsb_index_t *idx);
int64_t bsb_index_hbm_bytes(const bsb_index_t *idx);
int bsb_index_n_contigs(const bsb_index_t *idx);  /* includes the hidden crick copies */

/* replaces main_mem (fastmap.c:95-363): argv is exactly what BSBolt passes after the program name,
 * i.e. argv[0] == "mem", options, <idxbase> <in1.fq> [in2.fq]. SAM goes to out_fd, the log and the
 * `BSStat ...` lines go to log_fd. idx may be NULL (the index named in argv is loaded on device
 * `device` and freed at the end). Returns 0 on success. */
int bsb_mem_main(bsb_index_t *idx, int device, int argc, char **argv, int out_fd, int log_fd, bsb_run_stats_t *stats);

/* replaces the reference's whole output pipeline `bwa mem ... | stream_bam -@ <threads> -o <bam_path>`
 * (bsbolt/Align/AlignReads.py:52-60; bsbolt/External/HTSLIB/stream_bam.c): as bsb_mem_main, but the records are
 * encoded as BAM (what htslib's sam_parse1 + bam_write1 make of the same SAM lines: the uncompressed BAM stream is
 * byte-identical to the reference's) and BGZF-compressed:
 *   level -1 (default)  on the GPU: the read-group arbiter (samSorter, bs_sorter.cpp:84-170), the records and one dynamic-Huffman
 *                       deflate block per <= 0xff00 bytes (RFC 1951; BGZF framing of htslib's bgzf.c) are made by kernels and only
 *                       finished blocks cross PCIe; the host appends them to the file.
 *   level 0..9          by `threads` host threads (<= 0: this process's share of the cores) with zlib at that level; also the path
 *                       of runs whose records the device formatter does not make (-C comments, ALT contigs, -p, @SQ lines in -H). */
int bsb_mem_main_bam(bsb_index_t *idx, int device, int argc, char **argv, const char *bam_path, int threads, int level,
                     int log_fd, bsb_run_stats_t *stats);
/* The same two entry points over several GPUs of one box: idx[0] from bsb_index_load, idx[1..] its bsb_index_clone on the
 * other devices. The input is read and cut into the reference's batches ONCE (kt_pipeline step 0, fastmap.c:10-36), batch b
 * is aligned on idx[b mod n_idx] (the role of kt_for's worker threads, kthread.c:119-147 / bwamem.c:1319-1348: reads are
 * independent, nothing is exchanged between devices), and the records leave in input order (step 2, fastmap.c:61-73):
 * output and BSStat lines are byte-identical to the single-device run. */
int bsb_mem_main_multi(bsb_index_t *const *idx, int n_idx, int argc, char **argv, int out_fd, int log_fd, bsb_run_stats_t *stats);
int bsb_mem_main_multi_bam(bsb_index_t *const *idx, int n_idx, int argc, char **argv, const char *bam_path, int threads, int level,
                           int log_fd, bsb_run_stats_t *stats);

/* replaces stream_bam itself (HTSLIB/stream_bam.c main): SAM text on in_fd -> BAM file. Host only (needs no device).
 * Returns the number of records written, -1 on error. */
int64_t bsb_stream_bam(int in_fd, const char *bam_path, int threads, int level);

/* replaces bseq_read (bwa.c:73-145) for reads already in host memory: builds the bseq entries of one
 * batch (conversion-pattern assessment, undirectional duplication). r2 may be NULL (single end).
 * opt_argc/opt_argv: `bwa mem` options only (argv[0] == "mem", no positional arguments). */
bsb_batch_t *bsb_batch_create(bsb_index_t *idx, int opt_argc, char **opt_argv, int n, const bsb_read_t *r1, const bsb_read_t *r2);
/* replaces mem_process_seqs (bwamem.c:1319-1348): H2D, the eight kernels, D2H.
 * n_processed = bseq entries of all earlier batches (fastmap.c:59). */
int bsb_batch_align(bsb_batch_t *b, int64_t n_processed, bsb_run_stats_t *stats);
/* replaces step 2 of process() + samSorter (fastmap.c:61-73, bs_sorter.cpp): SAM text of the batch in
 * input order after the conversion-group arbitration; the buffer is owned by the batch. */
int bsb_batch_sam(bsb_batch_t *b, const char **sam, size_t *len, bsb_run_stats_t *stats);
int bsb_batch_n_entries(const bsb_batch_t *b);
void bsb_batch_free(bsb_batch_t *b);

/* replaces `bwa index -a bwtsw <fasta>` as run by `bsbolt Index` (bwa_idx_build, bwtindex.c:256-321;
 * bns_fasta2bntseq, bntseq.c:296-361): writes <prefix>.pac .opac .ann .amb .bwt .sa byte-identical to the
 * reference for references below 2^31 bases; the suffix array is built on the GPU. device_ms (optional)
 * receives the device time. */
int bsb_index_build(const char *fasta, const char *prefix, int device, double *device_ms);

/* Measurement aid for the seeding roofline: bandwidth of random 32-byte sector reads on `device` over a buffer of about
 * footprint_bytes (rounded to the nearest power of two; give the size of the occ-block array the kernel walks) -- `independent`:
 * every thread issues unrelated loads (what the memory system can deliver); `chase`: every thread's next address depends on
 * the sector it has just read, one load in flight per thread at full occupancy (the access pattern of backward search, the
 * honest ceiling for bwt_extend chains). GB/s; returns 0 on success. */
int bsb_random_sector_peak(int device, size_t footprint_bytes, double *gbs_independent, double *gbs_chase);

/* SAM header as printed by bwa_print_sam_hdr (bwa.c:530-553); buffer valid until the next call on this thread */
const char *bsb_sam_header(bsb_index_t *idx, int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
