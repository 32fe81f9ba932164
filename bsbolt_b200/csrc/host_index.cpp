// host_index.cpp -- see host_index.h
#include "host_index.h"
#include <algorithm>
#include <stdio.h>
#include <string.h>
#include <stdexcept>

namespace bsb {

static std::string infer_prefix(const std::string &hint)
{   // bwa_idx_infer_prefix (bwa.c:362-386)
    FILE *fp;
    std::string p = hint + ".64.bwt";
    if ((fp = fopen(p.c_str(), "rb")) != nullptr) { fclose(fp); return hint + ".64"; }
    p = hint + ".bwt";
    if ((fp = fopen(p.c_str(), "rb")) != nullptr) { fclose(fp); return hint; }
    throw std::runtime_error("[E::bsb_index_load] fail to locate the index files for '" + hint + "'");
}

static FILE *xopen_rb(const std::string &fn, const char *mode = "rb")
{
    FILE *fp = fopen(fn.c_str(), mode);
    if (!fp) throw std::runtime_error("[E::bsb_index_load] fail to open file '" + fn + "'");
    return fp;
}

static void xread(void *dst, size_t size, size_t n, FILE *fp, const std::string &fn)
{
    if (n && fread(dst, size, n, fp) != n) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] unexpected end of file in '" + fn + "'"); }
}

void HostIndex::load(const std::string &hint)
{
    prefix = infer_prefix(hint);
    { // .bwt : primary, L2[1..4], interleaved occ+bwt words
        std::string fn = prefix + ".bwt";
        FILE *fp = xopen_rb(fn);
        fseek(fp, 0, SEEK_END);
        long sz = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        if (sz < 40) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] truncated " + fn); }
        bwt_size = (uint64_t)(sz - 8 * 5) >> 2;
        xread(&primary, 8, 1, fp, fn);
        L2[0] = 0;
        xread(L2 + 1, 8, 4, fp, fn);
        bwt.assign(((bwt_size + 15) / 16 + 1) * 16, 0u); // whole blocks: every occ query reads 64 bytes
        xread(bwt.data(), 4, bwt_size, fp, fn);
        seq_len = L2[4];
        fclose(fp);
    }
    { // .sa : primary, L2[1..4], sa_intv, seq_len, n_sa-1 samples (sa[0] is implicit -1)
        std::string fn = prefix + ".sa";
        FILE *fp = xopen_rb(fn);
        uint64_t p2, skip[4], intv, sl;
        xread(&p2, 8, 1, fp, fn);
        xread(skip, 8, 4, fp, fn);
        xread(&intv, 8, 1, fp, fn);
        xread(&sl, 8, 1, fp, fn);
        if (p2 != primary || sl != seq_len) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] SA-BWT inconsistency in " + fn); }
        sa_intv = (int)intv;
        n_sa = (seq_len + sa_intv) / sa_intv;
        sa.assign(n_sa, 0);
        sa[0] = (uint64_t)-1;
        xread(sa.data() + 1, 8, n_sa - 1, fp, fn);
        fclose(fp);
    }
    std::vector<std::string> raw_names;   // as in the .ann file (the crick copies carry their _crick_bs suffix): what .alt is matched against
    { // .ann
        std::string fn = prefix + ".ann";
        FILE *fp = xopen_rb(fn, "r");
        long long xx; int n_seqs; unsigned seed;
        if (fscanf(fp, "%lld%d%u", &xx, &n_seqs, &seed) != 3) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] parse error reading " + fn); }
        l_pac = xx;
        contigs.resize(n_seqs);
        crick_l = 0;
        std::vector<char> str(8192);
        for (int i = 0; i < n_seqs; ++i) {
            HostContig &p = contigs[i];
            if (fscanf(fp, "%u%8191s", &p.gi, str.data()) != 2) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] parse error reading " + fn); }
            p.name = str.data();
            raw_names.push_back(p.name);
            p.is_crick = strstr(p.name.c_str(), "_crick_bs") != nullptr; // checkRname (bs_helpers.cpp:65-72)
            if (p.is_crick) p.name.resize(p.name.size() - 9);             // formatCrickRname (bs_helpers.cpp:74-76)
            std::string rest;
            int c;
            while ((c = fgetc(fp)) != '\n' && c != EOF) rest.push_back((char)c);
            if (c == EOF) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] unexpected end of file in " + fn); }
            if (rest.size() > 1 && rest != " (null)") p.anno = rest.substr(1);
            long long off;
            if (fscanf(fp, "%lld%d%d", &off, &p.len, &p.n_ambs) != 3) { fclose(fp); throw std::runtime_error("[E::bsb_index_load] parse error reading " + fn); }
            p.offset = off;
            p.is_alt = 0;
            if (!crick_l && p.is_crick) crick_l = p.offset;
        }
        fclose(fp);
    }
    { // .alt (optional): names of ALT contigs (bntseq.c:186-217)
        FILE *fp = fopen((prefix + ".alt").c_str(), "r");
        if (fp) {
            char line[4096];
            while (fgets(line, sizeof line, fp)) {
                if (line[0] == '@') continue;
                size_t l = strcspn(line, "\t\r\n");
                std::string nm(line, l);
                for (size_t k = contigs.size(); k-- > 0;)   // bns_restore hashes the untrimmed names; of equal names the last one wins
                    if (raw_names[k] == nm) { contigs[k].is_alt = 1; break; }
            }
            fclose(fp);
        }
    }
    { // .pac / .opac
        size_t nbytes = (size_t)(l_pac / 4 + 1);
        for (int k = 0; k < 2; ++k) {
            std::string fn = prefix + (k ? ".opac" : ".pac");
            FILE *fp = xopen_rb(fn);
            std::vector<uint8_t> &dst = k ? opac : pac;
            dst.assign(nbytes + 8, 0);
            xread(dst.data(), 1, nbytes, fp, fn);
            fclose(fp);
        }
    }
    anns.resize(contigs.size());
    for (size_t i = 0; i < contigs.size(); ++i) {
        anns[i].offset = contigs[i].offset; anns[i].len = contigs[i].len;
        anns[i].is_alt = contigs[i].is_alt; anns[i].is_crick = contigs[i].is_crick; anns[i].pad_ = 0;
    }
    build_sam_table();
}

void HostIndex::build_sam_table()
{
    const size_t n = contigs.size();
    ctg_text.clear(); ctg_name_off.assign(n + 1, 0); ctg_anno_off.assign(n + 1, 0); ctg_is_crick.assign(n, 0); ctg_sign.assign(n, 0);
    any_alt = false;
    for (size_t i = 0; i < n; ++i) {
        ctg_name_off[i] = (uint32_t)ctg_text.size();
        ctg_text.insert(ctg_text.end(), contigs[i].name.begin(), contigs[i].name.end());
        ctg_is_crick[i] = contigs[i].is_crick ? 1 : 0;
        ctg_sign[i] = (uint8_t)((contigs[i].name.find('+') != std::string::npos ? 1 : 0) | (contigs[i].name.find('-') != std::string::npos ? 2 : 0));
        if (contigs[i].is_alt) any_alt = true;
    }
    ctg_name_off[n] = (uint32_t)ctg_text.size();
    for (size_t i = 0; i < n; ++i) {
        ctg_anno_off[i] = (uint32_t)ctg_text.size();
        ctg_text.insert(ctg_text.end(), contigs[i].anno.begin(), contigs[i].anno.end());
    }
    ctg_anno_off[n] = (uint32_t)ctg_text.size();
    ctg_sorted.resize(n);
    for (size_t i = 0; i < n; ++i) ctg_sorted[i] = (int32_t)i;
    std::stable_sort(ctg_sorted.begin(), ctg_sorted.end(), [&](int32_t a, int32_t b) { return contigs[a].name < contigs[b].name; });
}

IndexView HostIndex::host_view() const
{
    IndexView v;
    v.bwt = bwt.data(); v.sa = sa.data(); v.pac = pac.data(); v.opac = opac.data(); v.anns = anns.data();
    v.primary = primary; for (int i = 0; i < 5; ++i) v.L2[i] = L2[i];
    v.seq_len = seq_len; v.l_pac = l_pac; v.crick_l = crick_l;
    v.n_seqs = (int)anns.size(); v.sa_intv = sa_intv;
    v.sa32 = nullptr; v.sa32_intv = 0; v.pad_ = 0; v.occ32 = nullptr; v.sa_hi = nullptr;
    return v;
}

} // namespace bsb
