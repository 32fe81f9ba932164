// host_index.h -- loads the index files written by `bsbolt Index` (reference `bwa index`) into host
// memory, ready for upload to HBM. File formats: SURVEY.md section 8 f-1;
// reference readers: bwt_restore_bwt / bwt_restore_sa (bwt.c:396-462), bns_restore_core (bntseq.c:99-167),
// bwa_idx_load_from_disk (bwa.c:407-438).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "bsb_types.h"
#include "bsb_index.h"

namespace bsb {

struct HostContig {
    std::string name, anno;
    int64_t offset; int32_t len, n_ambs, is_alt, is_crick; uint32_t gi;
};

struct HostIndex {
    uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0, bwt_size = 0;
    std::vector<uint32_t> bwt;      // occ-interleaved, bwt_size words (padded to whole 64-byte blocks)
    int sa_intv = 0; uint64_t n_sa = 0;
    std::vector<uint64_t> sa;
    int64_t l_pac = 0, crick_l = 0;
    std::vector<uint8_t> pac, opac; // l_pac/4+1 bytes each
    std::vector<HostContig> contigs;
    std::vector<Ann> anns;
    std::string prefix;
    // contig table flattened for the SAM formatter (bsb_sam.h): names back to back, then annotations
    std::vector<char> ctg_text; std::vector<uint32_t> ctg_name_off, ctg_anno_off; std::vector<uint8_t> ctg_is_crick, ctg_sign;
    std::vector<int32_t> ctg_sorted;   // contig ids ordered by name, then id: the BAM encoder's name -> tid lookup (BamContigs, bsb_bam.h)
    bool any_alt = false;
    void build_sam_table();

    // throws std::runtime_error on any missing/corrupt file
    void load(const std::string &hint);
    IndexView host_view() const;    // view over the host copies (used by the CPU unit tests only)
};

} // namespace bsb
