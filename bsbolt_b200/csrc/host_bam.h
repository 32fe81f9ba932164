// host_bam.h -- SAM text -> BAM (BGZF) on the host, multi-threaded.
//
// Replaces the second half of the reference's output pipeline, `... | stream_bam -@ T -o <prefix>.bam`
// (bsbolt/Align/AlignReads.py:52-60, bsbolt/External/HTSLIB/stream_bam.c): what htslib's sam_parse1 (sam.c:1924-2160)
// + bam_write1 (sam.c:661-735) + bam_hdr_write make of a SAM stream, i.e. the *uncompressed* BAM byte stream is
// identical to the reference's; the BGZF framing differs (blocks are cut per worker and never split a record).
// At the default level the pipeline does not come through records() at all: the aligner hands over BGZF blocks it compressed on
// the device (BatchResult::bam; bsb_bam.h, bsb_deflate.h) and blocks() appends them.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <unordered_map>
#include <vector>

namespace bsb {

class BamWriter {
public:
    // level: zlib level 0..9, -1 = zlib's default (what hts_open "wb" uses)
    BamWriter(const std::string &path, int threads, int level);
    ~BamWriter();
    void header(const std::string &sam_header_text);     // magic, header text, reference table from the @SQ lines
    void records(const char *text, size_t n);            // whole SAM record lines, '\n'-terminated
    void blocks(const uint8_t *bgzf, size_t n, uint64_t raw_bytes, uint64_t n_records);   // finished BGZF blocks (compressed on the device), appended as they are
    // the pipeline asks the aligner for finished blocks (BatchResult::want_bam) when this is set: the default compression level
    // (the caller did not ask for a zlib level) and no BSB_BAM_HOST in the environment
    void accept_device_blocks(bool yes) { device_blocks_ = yes; }
    bool device_blocks() const { return device_blocks_; }
    void close();                                        // flush + BGZF EOF block
    uint64_t n_records() const { return n_records_; }
    uint64_t raw_bytes() const { return raw_bytes_; }    // uncompressed BAM bytes written so far
    uint64_t file_bytes() const { return file_bytes_; }
    double sec_busy() const { return sec_busy_; }
    // one SAM line [p, e) (no newline) appended to `out` as a BAM record; throws std::runtime_error on malformed input
    void encode_record(const char *p, const char *e, std::vector<uint8_t> &out) const;

private:
    struct Worker;
    FILE *f_ = nullptr;
    int threads_, level_;
    bool have_header_ = false, closed_ = false, device_blocks_ = false;
    std::unordered_map<std::string, int> ref_ids_;
    std::vector<Worker *> workers_;
    uint64_t n_records_ = 0, raw_bytes_ = 0, file_bytes_ = 0;
    double sec_busy_ = 0;
    int tid_of(const char *p, size_t n) const;
};

// SAM stream on in_fd -> BAM file: the whole of stream_bam.c. Returns the number of records.
uint64_t stream_bam(int in_fd, const std::string &bam_path, int threads, int level);

} // namespace bsb
