// bamsim.cpp -- CPU harness for the device's BAM stage (TEST INFRASTRUCTURE, see bam_emulation.h).
//   bamsim deflate <in> <out.gz>   : the file as BGZF blocks through bgzf_block (bsb_deflate.h); every block is also inflated
//                                    with zlib here and its CRC-32 checked against zlib's
//   bamsim sam2bam <in.sam> <out>  : a SAM file as a BAM file the way the device makes it -- header by the host writer,
//                                    records through bam_entry with the sorted-contig lookup, blocks through bgzf_block
#include <stdio.h>
#include <string.h>
#include <zlib.h>
#include <fstream>
#include <iterator>
#include "bam_emulation.h"
#include "../../bsbolt_b200/csrc/host_bam.h"

using namespace bsb;

static std::vector<uint8_t> slurp(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// inflate every block with zlib and compare: contents, ISIZE and CRC-32
static void check_blocks(const std::vector<uint8_t> &bgzf, const uint8_t *raw, size_t n)
{
    size_t at = 0, done = 0;
    std::vector<uint8_t> buf(BGZF_MAX_IN + 16);
    long n_dyn = 0, n_stored = 0;
    while (at < bgzf.size()) {
        const uint8_t *b = bgzf.data() + at;
        if (at + 26 > bgzf.size() || b[0] != 0x1f || b[1] != 0x8b || b[12] != 'B' || b[13] != 'C') throw std::runtime_error("bad BGZF header");
        const size_t total = (size_t)(b[16] | b[17] << 8) + 1;
        if (at + total > bgzf.size()) throw std::runtime_error("BGZF block runs past the end");
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("inflateInit2");
        zs.next_in = const_cast<Bytef *>(b + 18); zs.avail_in = (uInt)(total - 26);
        zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
        const int rc = inflate(&zs, Z_FINISH);
        const size_t got = zs.total_out;
        const bool all_in = zs.avail_in == 0;
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) throw std::runtime_error("inflate failed on block at " + std::to_string(at) + ": rc " + std::to_string(rc) + (zs.msg ? std::string(" ") + zs.msg : ""));
        if (!all_in) throw std::runtime_error("deflate stream shorter than its block");
        if (done + got > n || memcmp(buf.data(), raw + done, got)) throw std::runtime_error("block at " + std::to_string(at) + " inflates to different bytes");
        const uint8_t *t = b + total - 8;
        const uint32_t crc = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
        const uint32_t isz = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
        if (isz != got) throw std::runtime_error("ISIZE mismatch");
        if (crc != (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw + done, (uInt)got)) throw std::runtime_error("CRC-32 mismatch on block at " + std::to_string(at));
        ((b[18] & 6) == 0 ? n_stored : n_dyn)++;
        done += got; at += total;
    }
    if (done != n) throw std::runtime_error("blocks cover " + std::to_string(done) + " of " + std::to_string(n) + " bytes");
    fprintf(stderr, "[bamsim] %zu bytes -> %zu bytes (%.3f), %ld dynamic + %ld stored blocks, all verified with zlib\n", n, bgzf.size(),
            n ? (double)bgzf.size() / (double)n : 0.0, n_dyn, n_stored);
}

static const uint8_t kEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

int main(int argc, char **argv)
{
    try {
        if (argc == 4 && !strcmp(argv[1], "deflate")) {
            const std::vector<uint8_t> raw = slurp(argv[2]);
            std::vector<uint8_t> out;
            bgzf_emulated(raw.data(), raw.size(), out);
            check_blocks(out, raw.data(), raw.size());
            FILE *f = fopen(argv[3], "wb");
            if (!f) throw std::runtime_error("cannot write the output");
            fwrite(out.data(), 1, out.size(), f);
            fwrite(kEof, 1, sizeof kEof, f);
            fclose(f);
            return 0;
        }
        if (argc == 4 && !strcmp(argv[1], "sam2bam")) {
            const std::vector<uint8_t> sam = slurp(argv[2]);
            const char *p = reinterpret_cast<const char *>(sam.data()), *end = p + sam.size();
            std::string header;
            while (p < end && *p == '@') { const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p)); const char *e = nl ? nl + 1 : end; header.append(p, e); p = e; }
            // contig table from the @SQ lines, flattened like HostIndex::build_sam_table
            std::vector<std::string> names;
            for (size_t q = 0; q < header.size();) {
                size_t e = header.find('\n', q);
                if (e == std::string::npos) e = header.size();
                if (header.compare(q, 4, "@SQ\t") == 0) {
                    const size_t s = header.find("SN:", q);
                    if (s != std::string::npos && s < e) { size_t t = header.find_first_of("\t\n", s); names.push_back(header.substr(s + 3, t - s - 3)); }
                }
                q = e + 1;
            }
            std::vector<char> ctg_text; std::vector<uint32_t> ctg_off(names.size() + 1); std::vector<int32_t> sorted(names.size());
            for (size_t i = 0; i < names.size(); ++i) { ctg_off[i] = (uint32_t)ctg_text.size(); ctg_text.insert(ctg_text.end(), names[i].begin(), names[i].end()); sorted[i] = (int32_t)i; }
            ctg_off[names.size()] = (uint32_t)ctg_text.size();
            std::stable_sort(sorted.begin(), sorted.end(), [&](int32_t a, int32_t b) { return names[a] < names[b]; });
            // one entry per line, all kept
            std::vector<uint32_t> text_off;
            const char *text = p;
            while (p < end) { text_off.push_back((uint32_t)(p - text)); const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p)); p = nl ? nl + 1 : end; }
            const int n = (int)text_off.size();
            text_off.push_back((uint32_t)(end - text));
            std::vector<uint8_t> code(n, BAM_KEEP);
            BamView v;
            memset(&v, 0, sizeof v);
            v.text = text; v.text_off = text_off.data(); v.code = code.data();
            v.ctg.text = ctg_text.data(); v.ctg.name_off = ctg_off.data(); v.ctg.sorted = sorted.data(); v.ctg.n = (int)names.size();
            std::vector<uint32_t> off(n + 1, 0);
            for (int i = 0; i < n; ++i) {
                SamCount c;
                const int rc = bam_entry(c, v, i, false);
                if (rc) throw std::runtime_error(std::string("[E::bam_encode] ") + bam_strerror(rc));
                off[i + 1] = off[i] + (uint32_t)c.n;
            }
            std::vector<uint8_t> raw(off[n] + 16);
            for (int i = 0; i < n; ++i) {
                SamWrite w = {reinterpret_cast<char *>(raw.data() + off[i])};
                if (bam_entry(w, v, i, true) || (uint64_t)(w.p - reinterpret_cast<char *>(raw.data())) != off[i + 1]) throw std::runtime_error("sizing and writing disagree");
            }
            std::vector<uint8_t> bgzf;
            bgzf_emulated(raw.data(), entry_cuts(off), bgzf);
            check_blocks(bgzf, raw.data(), (size_t)off[n]);
            BamWriter bw(argv[3], 1, 1);
            bw.header(header);
            bw.blocks(bgzf.data(), bgzf.size(), off[n], (uint64_t)n);
            bw.close();
            return 0;
        }
        fprintf(stderr, "usage: bamsim deflate <in> <out.gz> | bamsim sam2bam <in.sam> <out.bam>\n");
        return 1;
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
