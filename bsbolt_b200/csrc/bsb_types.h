// bsb_types.h -- plain-old-data records shared by the host pipeline and the sm_100a kernels.
//
// Every struct here is trivially copyable: the same bytes live in pinned host memory and in HBM.
// Field meanings follow the reference's records so that parity can be checked field by field:
//   Opt      <- mem_opt_t        (bwamem.h:26-62, defaults bwamem.c:50-88)
//   Intv     <- bwtintv_t        (bwt.h:60-62)
//   Seed     <- mem_seed_t       (bwamem.c:172-176)
//   Chain    <- mem_chain_t      (bwamem.c:178-184)
//   AlnReg   <- mem_alnreg_t     (bwamem.h:64-83)
//   PeStat   <- mem_pestat_t     (bwamem.h:87-91)
//   AlnOut   <- mem_aln_t        (bwamem.h:93-104) flattened for the host SAM writer
#pragma once
#include <stdint.h>

namespace bsb {

enum : int {
    F_PE = 0x2, F_NOPAIRING = 0x4, F_ALL = 0x8, F_NO_MULTI = 0x10, F_NO_RESCUE = 0x20,
    F_REF_HDR = 0x100, F_SOFTCLIP = 0x200, F_SMARTPE = 0x400, F_PRIMARY5 = 0x800,
    F_KEEP_SUPP_MAPQ = 0x1000, F_XB = 0x2000
};

struct Opt {
    int a, b, o_del, e_del, o_ins, e_ins;
    int pen_unpaired, pen_clip5, pen_clip3;
    int w, zdrop;
    uint64_t max_mem_intv;
    int T, flag, min_seed_len, min_chain_weight, max_chain_extend;
    float split_factor;
    int split_width, max_occ, max_chain_gap, n_threads, chunk_size;
    float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
    int mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
    int8_t mat[25];
    int undirectional, ch_conversion_threshold;
    float ch_conversion_proportion, substitution_proportion;
};

struct Intv {            // bi-directional SA interval
    uint64_t x0, x1, x2; // forward start, reverse-complement start, size
    uint64_t info;       // (query_begin << 32) | query_end
};

struct Seed {
    int64_t rbeg;
    int32_t qbeg, len;
    int32_t score;
    int32_t rid;         // contig id, <0: seed bridges contigs/strands and is dropped
};

struct Chain {
    int64_t pos;
    int32_t n;           // number of seeds
    int32_t head, tail;  // seed slots (linked through next[]) while chaining; head = offset into cseeds after compaction
    int32_t rid, first;
    int32_t w;
    int8_t  kept, is_alt;
    float   frac_rep;
};

struct AlnReg {
    int64_t rb, re;
    int32_t qb, qe;
    int32_t rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov;
    int32_t secondary, secondary_all, seedlen0;
    int32_t n_comp;
    int32_t is_alt;
    float   frac_rep;
    uint64_t hash;
};

struct PeStat {
    int32_t low, high, failed, pad_;
    double avg, std;
};

// One SAM line worth of alignment, produced on the device, formatted on the host.
struct AlnOut {
    int64_t pos;
    int32_t rid;        // <0: unmapped
    int32_t flag;
    int32_t is_rev, is_alt, mapq, NM;
    int32_t n_cigar;
    uint32_t cigar_off; // arena offsets (bytes); cigar = uint32[n_cigar]
    uint32_t md_off;    // "MD...\tXB:Z:..." without NUL; md_len bytes
    int32_t md_len;
    int32_t ch_meth, ch_unmeth, cg_meth, cg_unmeth;
    int32_t score, sub, alt_sc;
    uint32_t xa_off;    // XaOut[xa_n]
    int32_t xa_n;
};

struct XaOut {          // one XA:Z entry (mem_gen_alt, bwamem_extra.c:99-152)
    int64_t pos;
    int32_t rid, is_rev, NM, score, n_cigar;
    uint32_t cigar_off;
};

struct ReadOut {        // per bseq entry
    uint32_t aln_off;   // AlnOut[n_aln] in the arena
    int32_t n_aln;
    // the record mem_aln2sam receives as "mate" (h[] in mem_sam_pe): only the fields it reads
    int64_t h_pos;
    int32_t h_rid, h_is_rev, h_n_cigar, h_rlen, h_ch_meth, h_ch_unmeth;
    int32_t err;        // non-zero: scratch/arena overflow on this read (reported loudly by the host)
    int32_t pad_;
};

// Deferred alignment work: the finalisation kernel decides WHICH regions become records (flags, MAPQ,
// pairing); the CIGAR / NM / MD / XB of each chosen region is then produced by a separate, warp-cooperative
// kernel from this task list.
struct AlnTask {
    int64_t rb, re;      // reference span (doubled coordinates)
    int32_t qb, qe;      // query span
    int32_t truesc, w;   // local score of the region and the band it was found with (band inference)
    int32_t read;        // bseq entry (query bases)
    int32_t kind;        // 0: complete an AlnOut, 1: complete an XaOut
    uint32_t target_off; // arena offset of that record
    int32_t mate_read;   // >= 0: this alignment is also the "mate" record of that entry (h[] in mem_sam_pe)
};

// contig table entry (bntann1_t subset, bntseq.h:40-48)
struct Ann {
    int64_t offset;
    int32_t len, is_alt, is_crick, pad_;
};

enum : int {
    ERR_NONE = 0, ERR_INTV_OVERFLOW = 1, ERR_ARENA_OVERFLOW = 2, ERR_SCRATCH_OVERFLOW = 3,
    ERR_CIGAR_OVERFLOW = 4, ERR_NO_MD = 5,
    ERR_ROW_TILE = 6      // lane-per-read extension: a flank longer than the shared-memory row tile (the stage is re-run warp-per-read)
};

} // namespace bsb
