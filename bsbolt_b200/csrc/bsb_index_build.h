// bsb_index_build.h -- GPU index construction (see bsb_index_build.cu)
#pragma once
#include <stdint.h>
#include <string>

namespace bsb {
struct IndexBuildStats { int64_t l_pac = 0, seq_len = 0; int rounds = 0, n_contigs = 0; float ms_device = 0; };
// Writes <prefix>.pac .opac .ann .amb .bwt .sa, byte-identical to the reference's `bwa index -a bwtsw <fasta>`.
void index_build(const std::string &fasta, const std::string &prefix, int device, IndexBuildStats *stats);
}
