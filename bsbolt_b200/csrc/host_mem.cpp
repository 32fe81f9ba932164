// host_mem.cpp -- see host_mem.h
#include "host_mem.h"
#include <stdarg.h>
#include "host_bam.h"
#include "bsb_hd.h"
#include "bsb_sam.h"
#include <ctype.h>
#include <getopt.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <memory>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace bsb {

// ------------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------------
void fill_scmat(int a, int b, int8_t mat[25])
{
    int i, j, k;
    for (i = k = 0; i < 4; ++i) {
        for (j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = -1;
    }
    for (j = 0; j < 5; ++j) mat[k++] = -1;
}

void opt_init(Opt &o)
{
    memset(&o, 0, sizeof(Opt));
    o.a = 1; o.b = 4;
    o.o_del = o.o_ins = 6;
    o.e_del = o.e_ins = 1;
    o.w = 100;
    o.T = 30;
    o.zdrop = 100;
    o.pen_unpaired = 17;
    o.pen_clip5 = o.pen_clip3 = 5;
    o.max_mem_intv = 20;
    o.min_seed_len = 19;
    o.split_width = 10;
    o.max_occ = 500;
    o.max_chain_gap = 10000;
    o.max_ins = 10000;
    o.mask_level = 0.50f;
    o.drop_ratio = 0.50f;
    o.XA_drop_ratio = 0.95f;
    o.split_factor = 1.5f;
    o.chunk_size = 10000000;
    o.n_threads = 1;
    o.max_XA_hits = 5;
    o.max_XA_hits_alt = 200;
    o.max_matesw = 50;
    o.mask_level_redun = 0.95f;
    o.min_chain_weight = 0;
    o.max_chain_extend = 1 << 30;
    o.mapQ_coef_len = 50; o.mapQ_coef_fac = (int)log(o.mapQ_coef_len);
    o.undirectional = 0;
    o.ch_conversion_threshold = 5;
    o.ch_conversion_proportion = 0.5f;
    o.substitution_proportion = 0.1f;
    fill_scmat(o.a, o.b, o.mat);
}

static std::string unescape(const std::string &s)
{   // bwa_escape (bwa.c:555-571)
    std::string q;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '\\') {
            ++i;
            if (i >= s.size()) break;
            if (s[i] == 't') q.push_back('\t');
            else if (s[i] == 'n') q.push_back('\n');
            else if (s[i] == 'r') q.push_back('\r');
            else if (s[i] == '\\') q.push_back('\\');
        } else q.push_back(s[i]);
    }
    return q;
}

static void insert_header(MemArgs &ma, const std::string &s)
{   // bwa_insert_header (bwa.c:603-616)
    if (s.empty() || s[0] != '@') return;
    if (ma.have_hdr) { ma.hdr_line.push_back('\n'); ma.hdr_line += unescape(s); }
    else { ma.hdr_line = unescape(s); ma.have_hdr = true; }
}

static void parse_two(const char *arg, int *a, int *b)
{
    char *p;
    *a = *b = (int)strtol(arg, &p, 10);
    if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) *b = (int)strtol(p + 1, &p, 10);
}

static std::mutex g_getopt_mutex;

int parse_mem_args(int argc, char **argv, MemArgs &ma, std::string &err)
{
    std::lock_guard<std::mutex> lock(g_getopt_mutex);
    Opt &opt = ma.opt;
    Opt opt0;
    opt_init(opt);
    memset(&opt0, 0, sizeof(Opt));
    memset(ma.pes0, 0, sizeof(ma.pes0));
    for (int i = 0; i < 4; ++i) ma.pes0[i].failed = 1;
    std::string rg_line;
    const char *mode = nullptr;
    int c;
    // optind = 0 makes glibc re-initialise its scanner completely: with optind = 1 it keeps a pointer into the PREVIOUS call's
    // argv (freed by now when the caller is the Python layer), and whatever lies there is parsed as clustered option letters
    optind = 0; opterr = 0;
    while ((c = getopt(argc, argv, "51qpaMCSPVYjuzk:c:v:s:r:t:R:A:B:O:E:U:w:L:d:T:Q:D:m:I:N:o:f:W:x:G:h:Z:y:K:X:H:l:n:e:")) >= 0) {
        if (c == 'k') opt.min_seed_len = atoi(optarg), opt0.min_seed_len = 1;
        else if (c == '1') ma.no_mt_io = true;
        else if (c == 'x') mode = optarg;
        else if (c == 'w') opt.w = atoi(optarg), opt0.w = 1;
        else if (c == 'A') opt.a = atoi(optarg), opt0.a = 1;
        else if (c == 'B') opt.b = atoi(optarg), opt0.b = 1;
        else if (c == 'T') opt.T = atoi(optarg), opt0.T = 1;
        else if (c == 'U') opt.pen_unpaired = atoi(optarg), opt0.pen_unpaired = 1;
        else if (c == 't') opt.n_threads = atoi(optarg), opt.n_threads = opt.n_threads > 1 ? opt.n_threads : 1;
        else if (c == 'P') opt.flag |= F_NOPAIRING;
        else if (c == 'a') opt.flag |= F_ALL;
        else if (c == 'p') opt.flag |= F_PE | F_SMARTPE;
        else if (c == 'M') opt.flag |= F_NO_MULTI;
        else if (c == 'S') opt.flag |= F_NO_RESCUE;
        else if (c == 'Y') opt.flag |= F_SOFTCLIP;
        else if (c == 'V') opt.flag |= F_REF_HDR;
        else if (c == '5') opt.flag |= F_PRIMARY5 | F_KEEP_SUPP_MAPQ;
        else if (c == 'q') opt.flag |= F_KEEP_SUPP_MAPQ;
        else if (c == 'u') opt.flag |= F_XB;
        else if (c == 'z') opt.undirectional = 1;
        else if (c == 'e') opt.substitution_proportion = (float)atof(optarg);
        else if (c == 'c') opt.max_occ = atoi(optarg), opt0.max_occ = 1;
        else if (c == 'd') opt.zdrop = atoi(optarg), opt0.zdrop = 1;
        else if (c == 'v') ma.verbose = atoi(optarg);
        else if (c == 'j') ma.ignore_alt = true;
        else if (c == 'r') opt.split_factor = (float)atof(optarg), opt0.split_factor = 1.f;
        else if (c == 'D') opt.drop_ratio = (float)atof(optarg), opt0.drop_ratio = 1.f;
        else if (c == 'm') opt.max_matesw = atoi(optarg), opt0.max_matesw = 1;
        else if (c == 's') opt.split_width = atoi(optarg), opt0.split_width = 1;
        else if (c == 'G') opt.max_chain_gap = atoi(optarg), opt0.max_chain_gap = 1;
        else if (c == 'N') opt.max_chain_extend = atoi(optarg), opt0.max_chain_extend = 1;
        else if (c == 'o' || c == 'f') ma.out_path = optarg;
        else if (c == 'W') opt.min_chain_weight = atoi(optarg), opt0.min_chain_weight = 1;
        else if (c == 'y') opt.max_mem_intv = (uint64_t)atol(optarg), opt0.max_mem_intv = 1;
        else if (c == 'C') ma.copy_comment = true;
        else if (c == 'K') ma.fixed_chunk_size = atoi(optarg);
        else if (c == 'X') opt.mask_level = (float)atof(optarg);
        else if (c == 'Z') opt.XA_drop_ratio = (float)atof(optarg);
        else if (c == 'l') opt.ch_conversion_proportion = (float)atof(optarg);
        else if (c == 'n') opt.ch_conversion_threshold = atoi(optarg);
        else if (c == 'h') { opt0.max_XA_hits = opt0.max_XA_hits_alt = 1; parse_two(optarg, &opt.max_XA_hits, &opt.max_XA_hits_alt); }
        else if (c == 'Q') {
            opt0.mapQ_coef_len = 1;
            opt.mapQ_coef_len = (float)atoi(optarg);
            opt.mapQ_coef_fac = opt.mapQ_coef_len > 0 ? (int)log(opt.mapQ_coef_len) : 0;
        }
        else if (c == 'O') { opt0.o_del = opt0.o_ins = 1; parse_two(optarg, &opt.o_del, &opt.o_ins); }
        else if (c == 'E') { opt0.e_del = opt0.e_ins = 1; parse_two(optarg, &opt.e_del, &opt.e_ins); }
        else if (c == 'L') { opt0.pen_clip5 = opt0.pen_clip3 = 1; parse_two(optarg, &opt.pen_clip5, &opt.pen_clip3); }
        else if (c == 'R') {
            std::string s = optarg; // bwa_set_rg (bwa.c:573-601)
            if (s.compare(0, 3, "@RG") != 0) { err = "[E::bwa_set_rg] the read group line is not started with @RG"; return 1; }
            if (s.find('\t') != std::string::npos) { err = "[E::bwa_set_rg] the read group line contained literal <tab> characters -- replace with escaped tabs: \\t"; return 1; }
            rg_line = unescape(s);
            size_t p = rg_line.find("\tID:");
            if (p == std::string::npos) { err = "[E::bwa_set_rg] no ID within the read group line"; return 1; }
            p += 4;
            size_t q = p;
            while (q < rg_line.size() && rg_line[q] != '\t' && rg_line[q] != '\n') ++q;
            if (q - p + 1 > 256) { err = "[E::bwa_set_rg] @RG:ID is longer than 255 characters"; return 1; }
            ma.rg_id = rg_line.substr(p, q - p);
        }
        else if (c == 'H') {
            if (optarg[0] != '@') {
                FILE *fp;
                if ((fp = fopen(optarg, "r")) != nullptr) {
                    std::vector<char> buf(0x10000);
                    while (fgets(buf.data(), 0xffff, fp)) {
                        size_t l = strlen(buf.data());
                        if (l && buf[l - 1] == '\n') buf[l - 1] = 0;
                        insert_header(ma, buf.data());
                    }
                    fclose(fp);
                }
            } else insert_header(ma, optarg);
        }
        else if (c == 'I') {
            char *p;
            PeStat *pes = ma.pes0;
            ma.have_pes0 = true;
            pes[1].failed = 0;
            pes[1].avg = strtod(optarg, &p);
            pes[1].std = pes[1].avg * .1;
            if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].std = strtod(p + 1, &p);
            pes[1].high = (int)(pes[1].avg + 4. * pes[1].std + .499);
            pes[1].low = (int)(pes[1].avg - 4. * pes[1].std + .499);
            if (pes[1].low < 1) pes[1].low = 1;
            if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].high = (int)(strtod(p + 1, &p) + .499);
            if (*p != 0 && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].low = (int)(strtod(p + 1, &p) + .499);
        }
        else { err = "[E::main_mem] unrecognized option"; return 1; }
    }
    if (!rg_line.empty()) insert_header(ma, rg_line);
    if (opt.n_threads < 1) opt.n_threads = 1;
    if (optind + 1 >= argc || optind + 3 < argc) {
        err = "Usage: bwa mem [options] <idxbase> <in1.fq> [in2.fq]";
        return 1;
    }
    if (mode) {
        std::string m = mode;
        if (m == "intractg") {
            if (!opt0.o_del) opt.o_del = 16;
            if (!opt0.o_ins) opt.o_ins = 16;
            if (!opt0.b) opt.b = 9;
            if (!opt0.pen_clip5) opt.pen_clip5 = 5;
            if (!opt0.pen_clip3) opt.pen_clip3 = 5;
        } else if (m == "pacbio" || m == "pbref" || m == "ont2d") {
            if (!opt0.o_del) opt.o_del = 1;
            if (!opt0.e_del) opt.e_del = 1;
            if (!opt0.o_ins) opt.o_ins = 1;
            if (!opt0.e_ins) opt.e_ins = 1;
            if (!opt0.b) opt.b = 1;
            if (opt0.split_factor == 0.) opt.split_factor = 10.;
            if (m == "ont2d") {
                if (!opt0.min_chain_weight) opt.min_chain_weight = 20;
                if (!opt0.min_seed_len) opt.min_seed_len = 14;
            } else {
                if (!opt0.min_chain_weight) opt.min_chain_weight = 40;
                if (!opt0.min_seed_len) opt.min_seed_len = 17;
            }
            if (!opt0.pen_clip5) opt.pen_clip5 = 0;
            if (!opt0.pen_clip3) opt.pen_clip3 = 0;
        } else { err = std::string("[E::main_mem] unknown read type '") + mode + "'"; return 1; }
    } else if (opt0.a) { // update_a (fastmap.c:78-92)
        if (!opt0.b) opt.b *= opt.a;
        if (!opt0.T) opt.T *= opt.a;
        if (!opt0.o_del) opt.o_del *= opt.a;
        if (!opt0.e_del) opt.e_del *= opt.a;
        if (!opt0.o_ins) opt.o_ins *= opt.a;
        if (!opt0.e_ins) opt.e_ins *= opt.a;
        if (!opt0.zdrop) opt.zdrop *= opt.a;
        if (!opt0.pen_clip5) opt.pen_clip5 *= opt.a;
        if (!opt0.pen_clip3) opt.pen_clip3 *= opt.a;
        if (!opt0.pen_unpaired) opt.pen_unpaired *= opt.a;
    }
    fill_scmat(opt.a, opt.b, opt.mat);
    ma.idxbase = argv[optind];
    ma.fq1 = argv[optind + 1];
    if (optind + 2 < argc) {
        if (opt.flag & F_PE) {
            if (ma.verbose >= 2) fprintf(stderr, "[W::%s] when '-p' is in use, the second query file is ignored.\n", "main_mem");
        } else {
            ma.fq2 = argv[optind + 2];
            opt.flag |= F_PE;
        }
    }
    if (const char *e = getenv("BSB_SHARD_COUNT")) ma.shard_count = atoi(e) > 1 ? atoi(e) : 1;
    if (const char *e = getenv("BSB_SHARD_INDEX")) ma.shard_index = atoi(e);
    if (const char *e = getenv("BSB_SHARD_PARTS")) ma.shard_parts = e;
    if (ma.shard_index < 0 || ma.shard_index >= ma.shard_count) { err = "[E::main_mem] BSB_SHARD_INDEX out of range"; return 1; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// SAM text
// ------------------------------------------------------------------------------------------------
std::string sam_header(const HostIndex &idx, const MemArgs &ma)
{
    std::string h;
    int n_SQ = 0;
    if (ma.have_hdr) {
        size_t p = 0;
        while ((p = ma.hdr_line.find("@SQ\t", p)) != std::string::npos) {
            if (p == 0 || ma.hdr_line[p - 1] == '\n') ++n_SQ;
            p += 4;
        }
    }
    if (n_SQ == 0) {
        for (const HostContig &c : idx.contigs) {
            if (c.is_crick) continue;
            h += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len);
            if (c.is_alt && !ma.ignore_alt) h += "\tAH:*";
            h += '\n';
        }
    }
    if (ma.have_hdr) { h += ma.hdr_line; h += '\n'; }
    if (!ma.pg_line.empty()) { h += ma.pg_line; h += '\n'; }
    return h;
}

// the text sink of the host formatter
struct SamString {
    std::string &s;
    void ch(char c) { s.push_back(c); }
    void mem(const char *p, size_t l) { s.append(p, l); }
    void num(long v)
    {
        char buf[24]; int l = 0; unsigned long x = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
        do { buf[l++] = (char)('0' + x % 10); x /= 10; } while (x);
        if (v < 0) buf[l++] = '-';
        while (l) s.push_back(buf[--l]);
    }
    void pa(double r) { char buf[64]; snprintf(buf, sizeof buf, "\tpa:f:%.3f", r); s += buf; }
};

static const unsigned char kNt4[256] = {
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};

// everything bsb_sam.h reads, over the host copies of one batch and its results
SamView make_sam_view(const MemArgs &ma, const HostIndex &idx, const ReadBatch &b, const BatchResult &res)
{
    SamView v;
    memset(&v, 0, sizeof v);
    v.flag = ma.opt.flag; v.ch_conversion_threshold = ma.opt.ch_conversion_threshold; v.ch_conversion_proportion = ma.opt.ch_conversion_proportion;
    v.rg_id = ma.rg_id.data(); v.rg_len = (int)ma.rg_id.size();
    v.names = b.names.data(); v.name_off = b.name_off.data();
    v.bases = b.bases.data(); v.qual = b.qual.data(); v.seq_off = b.seq_off.data();
    v.has_qual = b.has_qual.data(); v.pattern = b.pattern.data();
    v.cmt = b.comments.data(); v.cmt_off = b.cmt_off.data();
    v.ctg_text = idx.ctg_text.data(); v.ctg_name_off = idx.ctg_name_off.data(); v.ctg_anno_off = idx.ctg_anno_off.data();
    v.ctg_is_crick = idx.ctg_is_crick.data(); v.ctg_sign = idx.ctg_sign.data();
    v.arena = res.arena.data(); v.reads = res.reads.data();
    return v;
}

static void stats_from(const SamStats &s, EntryStats &st)
{
    st.alignment_score = s.alignment_score; st.mapped = s.mapped; st.bs_conflict = s.bs_conflict; st.crick = s.crick; st.paired = s.paired;
}

void format_entry(const MemArgs &ma, const HostIndex &idx, const ReadBatch &b, int i, const BatchResult &res,
                  std::string &sam, EntryStats &st)
{
    sam.clear();
    const SamView v = make_sam_view(ma, idx, b, res);
    SamString o = {sam};
    SamStats ss;
    sam_entry(o, v, i, (ma.opt.flag & F_PE) != 0, ss);
    stats_from(ss, st);
}

// the same, appended to a buffer shared by consecutive entries (the pipeline's formatter)
static void format_entry_append(const SamView &v, bool is_pe, int i, std::string &buf, EntryStats &st)
{
    SamString o = {buf};
    SamStats ss;
    sam_entry(o, v, i, is_pe, ss);
    stats_from(ss, st);
}

// ------------------------------------------------------------------------------------------------
// read-group arbiter
// ------------------------------------------------------------------------------------------------
void MapStats::add(const MapStats &o)
{
    reads += o.reads; alignments += o.alignments; wc2t += o.wc2t; wg2a += o.wg2a; cc2t += o.cc2t; cg2a += o.cg2a;
    unaligned += o.unaligned; bs_ambiguous += o.bs_ambiguous;
}

static void set_unmapped(const ReadBatch &b, int i, const EntryStats &st, std::string &sam)
{   // samSorter::setUnmapped (bs_sorter.cpp:51-82)
    sam.clear();
    sam.append(b.names.data() + b.name_off[i], b.name_off[i + 1] - b.name_off[i]);
    sam.push_back('\t');
    if (st.paired) sam += b.first[i] ? "77\t" : "141\t";
    else sam += "4\t";
    sam += "*\t0\t0\t*\t*\t0\t0\t";
    const char *bases = b.bases.data() + b.seq_off[i];
    int l = b.len(i);
    for (int k = 0; k < l; ++k) sam.push_back("ACGTN"[kNt4[(unsigned char)bases[k]]]);
    sam.push_back('\t');
    if (b.has_qual[i]) sam.append(b.qual.data() + b.seq_off[i], l);
    sam += "\tAS:i:0\tYS:Z:WC\n";
}

// Decision pass of the arbiter: which entries are printed (in input order), with BS-ambiguous groups rewritten
// as unmapped. Serial and cheap -- no SAM text is copied here.
static void sam_sort_range(const ReadBatch &b, std::vector<EntryStats> &st, int lo, int hi, std::vector<int> &emit, std::vector<uint8_t> &rewrite, MapStats &ms);

static void sam_sort_plan(const ReadBatch &b, std::vector<EntryStats> &st, std::vector<int> &emit, std::vector<uint8_t> &rewrite, MapStats &ms)
{
    emit.clear();
    rewrite.assign(b.n, 0);
    if (b.n == 0) return;
    sam_sort_range(b, st, 0, b.n, emit, rewrite, ms);
}

// The same over several threads: the entries are cut where a new read name starts (a group never straddles two ranges), every
// range is arbitrated on its own, the results are joined in input order.
static void sam_sort_plan_parallel(const ReadBatch &b, std::vector<EntryStats> &st, std::vector<int> &emit, std::vector<uint8_t> &rewrite, MapStats &ms,
                                   int n_threads)
{
    if (n_threads <= 1 || b.n < 4096) { sam_sort_plan(b, st, emit, rewrite, ms); return; }
    auto same_name = [&](int i, int j) {
        uint32_t li = b.name_off[i + 1] - b.name_off[i], lj = b.name_off[j + 1] - b.name_off[j];
        return li == lj && memcmp(b.names.data() + b.name_off[i], b.names.data() + b.name_off[j], li) == 0;
    };
    std::vector<int> cut(n_threads + 1, b.n);
    cut[0] = 0;
    for (int t = 1; t < n_threads; ++t) {
        int c = (int)((int64_t)b.n * t / n_threads);
        if (c < cut[t - 1]) c = cut[t - 1];
        while (c > 0 && c < b.n && same_name(c, c - 1)) ++c;   // forward to the first entry of the next name group
        cut[t] = c;
    }
    rewrite.assign(b.n, 0);
    std::vector<std::vector<int>> part(n_threads);
    std::vector<MapStats> pms(n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t] { if (cut[t] < cut[t + 1]) sam_sort_range(b, st, cut[t], cut[t + 1], part[t], rewrite, pms[t]); });
    for (auto &x : th) x.join();
    emit.clear();
    for (int t = 0; t < n_threads; ++t) { emit.insert(emit.end(), part[t].begin(), part[t].end()); ms.add(pms[t]); }
}

static void sam_sort_range(const ReadBatch &b, std::vector<EntryStats> &st, int lo, int hi, std::vector<int> &emit, std::vector<uint8_t> &rewrite, MapStats &ms)
{
    emit.clear();
    std::vector<int> g[2];
    long score[2] = {0, 0};
    auto same_name = [&](int i, int j) {
        uint32_t li = b.name_off[i + 1] - b.name_off[i], lj = b.name_off[j + 1] - b.name_off[j];
        return li == lj && memcmp(b.names.data() + b.name_off[i], b.names.data() + b.name_off[j], li) == 0;
    };
    auto update = [&](int i) {
        ++ms.alignments;
        if (st[i].bs_conflict) ++ms.bs_ambiguous;
        switch (st[i].mapped) {
            case 0: ++ms.unaligned; break;
            case 1: ++ms.cg2a; break;
            case 2: ++ms.cc2t; break;
            case 3: ++ms.wg2a; break;
            case 4: ++ms.wc2t; break;
        }
    };
    auto flush = [&]() {
        int pick = score[0] > score[1] ? 0 : score[0] < score[1] ? 1 : 2;
        ++ms.reads;
        if (pick == 1) {
            for (int i : g[1]) { update(i); emit.push_back(i); }
        } else {
            for (int i : g[0]) {
                if (pick == 2) {
                    if (st[i].mapped) rewrite[i] = 1;     // printed as an unmapped record (set_unmapped)
                    st[i].mapped = 0;
                    st[i].bs_conflict = 1;
                }
                update(i);
                emit.push_back(i);
            }
        }
        g[0].clear(); g[1].clear(); score[0] = score[1] = 0;
    };
    auto bank = [&](int i) { int k = b.read_group[i] ? 1 : 0; score[k] += st[i].alignment_score; g[k].push_back(i); };
    int cur = lo;
    bank(lo);
    for (int i = lo + 1; i < hi; ++i) {
        if (same_name(i, cur)) bank(i);
        else { flush(); cur = i; bank(i); }
    }
    flush();
}

void sam_sort_batch(const ReadBatch &b, std::vector<std::string> &sam, std::vector<EntryStats> &st, std::string &out, MapStats &ms)
{
    std::vector<int> emit;
    std::vector<uint8_t> rewrite;
    sam_sort_plan(b, st, emit, rewrite, ms);
    for (int i : emit) {
        if (rewrite[i]) set_unmapped(b, i, st[i], sam[i]);
        out += sam[i];
    }
}

// ------------------------------------------------------------------------------------------------
// insert size statistics + math tables
// ------------------------------------------------------------------------------------------------
void estimate_pestat(const Opt &opt, const std::vector<int8_t> &dir, const std::vector<int64_t> &isz, PeStat pes[4], int verbose, std::string *log_text)
{
    // the messages of one batch leave as ONE block (through the caller's log under its lock, or stderr): device slots run
    // concurrently and a line must never land inside another thread's line
    std::string text;
    auto say = [&](const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        text += buf;
    };
    struct Flush { std::string &t; std::string *sink; ~Flush() { if (sink) *sink += t; else if (!t.empty()) fputs(t.c_str(), stderr); } } flush = {text, log_text};
    (void)opt;
    std::vector<uint64_t> isize[4];
    memset(pes, 0, 4 * sizeof(PeStat));
    for (size_t i = 0; i < dir.size(); ++i)
        if (dir[i] >= 0) isize[dir[i]].push_back((uint64_t)isz[i]);
    if (verbose >= 3)
        say("[M::%s] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", "mem_pestat",
                (long)isize[0].size(), (long)isize[1].size(), (long)isize[2].size(), (long)isize[3].size());
    for (int d = 0; d < 4; ++d) {
        PeStat *r = &pes[d];
        std::vector<uint64_t> &q = isize[d];
        int p25, p50, p75, x;
        size_t n = q.size(), i;
        if (n < 10) {
            if (verbose >= 3) say("[M::%s] skip orientation %c%c as there are not enough pairs\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
            r->failed = 1;
            continue;
        } else if (verbose >= 3) say("[M::%s] analyzing insert size distribution for orientation %c%c...\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
        { // insert sizes are bounded by max_ins: counting sort when the range is small, else a comparison sort
            uint64_t mx = 0;
            for (uint64_t v : q) mx = v > mx ? v : mx;
            if (mx < (1u << 20)) {
                std::vector<uint32_t> cnt(mx + 1, 0);
                for (uint64_t v : q) ++cnt[v];
                size_t k = 0;
                for (uint64_t v = 0; v <= mx; ++v) for (uint32_t c = cnt[v]; c; --c) q[k++] = v;
            } else std::sort(q.begin(), q.end());
        }
        p25 = (int)q[(int)(.25 * n + .499)];
        p50 = (int)q[(int)(.50 * n + .499)];
        p75 = (int)q[(int)(.75 * n + .499)];
        r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
        if (r->low < 1) r->low = 1;
        r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
        if (verbose >= 3) {
            say("[M::%s] (25, 50, 75) percentile: (%d, %d, %d)\n", "mem_pestat", p25, p50, p75);
            say("[M::%s] low and high boundaries for computing mean and std.dev: (%d, %d)\n", "mem_pestat", r->low, r->high);
        }
        for (i = 0, x = 0, r->avg = 0; i < n; ++i)
            if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) r->avg += q[i], ++x;
        r->avg /= x;
        for (i = 0, r->std = 0; i < n; ++i)
            if (q[i] >= (uint64_t)r->low && q[i] <= (uint64_t)r->high) r->std += (q[i] - r->avg) * (q[i] - r->avg);
        r->std = sqrt(r->std / x);
        if (verbose >= 3) say("[M::%s] mean and std.dev: (%.2f, %.2f)\n", "mem_pestat", r->avg, r->std);
        r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
        r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
        if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
        if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
        if (r->low < 1) r->low = 1;
        if (verbose >= 3) say("[M::%s] low and high boundaries for proper pairs: (%d, %d)\n", "mem_pestat", r->low, r->high);
    }
    size_t max = 0;
    for (int d = 0; d < 4; ++d) max = max > isize[d].size() ? max : isize[d].size();
    for (int d = 0; d < 4; ++d)
        if (pes[d].failed == 0 && isize[d].size() < max * 0.05) {
            pes[d].failed = 1;
            if (verbose >= 3) say("[M::%s] skip orientation %c%c\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
        }
}

void build_log_table(std::vector<double> &t, int n)
{
    t.resize(n);
    for (int i = 0; i < n; ++i) t[i] = log((double)i); // integer arguments are converted to double by the C call
}

void build_pair_table(const Opt &opt, const PeStat pes[4], std::vector<double> &t, int off[4])
{
    t.clear();
    for (int d = 0; d < 4; ++d) {
        off[d] = (int)t.size();
        if (pes[d].failed || pes[d].high < pes[d].low) continue;
        for (int64_t dist = pes[d].low; dist <= pes[d].high; ++dist) {
            double ns = (dist - pes[d].avg) / pes[d].std;
            t.push_back(.721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt.a);
        }
    }
    if (t.empty()) t.push_back(0.);
}

// ------------------------------------------------------------------------------------------------
// the run
// ------------------------------------------------------------------------------------------------
static double now_sec()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

template <class F>
static void parallel_for(int n_threads, int n, F f, int serial_below = 256)
{
    if (n_threads <= 1 || n < serial_below) { for (int i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> th;
    int chunk = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        int lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=]() { for (int i = lo; i < hi; ++i) f(i); });
    }
    for (auto &x : th) x.join();
}

namespace {
// bounded single-producer/single-consumer hand-off between the pipeline stages
template <class T>
class Channel {
public:
    explicit Channel(size_t cap) : cap_(cap) {}
    void push(T v)
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return q_.size() < cap_ || closed_; });
        q_.push_back(std::move(v));
        cv_.notify_all();
    }
    bool pop(T &v)
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        cv_.notify_all();
        return true;
    }
    void close()
    {
        std::lock_guard<std::mutex> l(m_);
        closed_ = true;
        cv_.notify_all();
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

struct Job {
    BatchPlan plan;
    ReadBatch batch;
    BatchResult res;
    int64_t n_processed = 0;
    long batch_id = 0;
    long seq = 0;              // position in this run's stream of batches (batch_id skips other shards' batches)
    double sec_align = 0;
};

// hands finished batches to the formatter in input order, whichever device slot finished first
class OrderedDone {
public:
    void push(std::unique_ptr<Job> j)
    {
        std::lock_guard<std::mutex> l(m_);
        const long k = j->seq;
        ready_[k] = std::move(j);
        cv_.notify_all();
    }
    bool pop(std::unique_ptr<Job> &j)
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return aborted_ || ready_.count(next_) || (producers_ == 0 && ready_.empty()); });
        if (aborted_ || !ready_.count(next_)) return false;
        j = std::move(ready_[next_]);
        ready_.erase(next_);
        ++next_;
        return true;
    }
    void add_producer() { std::lock_guard<std::mutex> l(m_); ++producers_; }
    void producer_done() { std::lock_guard<std::mutex> l(m_); --producers_; cv_.notify_all(); }
    void abort() { std::lock_guard<std::mutex> l(m_); aborted_ = true; cv_.notify_all(); }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::map<long, std::unique_ptr<Job>> ready_;
    long next_ = 0;
    int producers_ = 0;
    bool aborted_ = false;
};
} // namespace

// Smart pairing (-p): process() with MEM_F_SMARTPE (fastmap.c:38-57). The batch was cut from ONE interleaved file
// (every read converted as a first mate); bseq_classify (bwa.c:147-165) splits it by read name into single-end
// entries and pairs of consecutive equal names, the two classes are aligned as separate calls -- single-end at
// n_processed, pairs at n_processed + n_single with their own insert-size statistics -- and only the SAM text
// is copied back to the entries (the classes are struct copies in the reference), so the arbiter sees every
// entry with alignment_score 0 / mapped 0. The result is handed on as device-style text (have_text).
static void align_smart_pairs(const MemArgs &ma, const HostIndex &idx, BatchAligner &aligner, const ReadBatch &b, int64_t n_processed,
                              BatchResult &res, int slot, FILE *log, std::mutex &log_m)
{
    std::vector<int> sep[2];
    auto same_name = [&](int x, int y) {   // strcmp: a name ends at its first NUL
        const char *px = b.names.data() + b.name_off[x], *py = b.names.data() + b.name_off[y];
        const size_t lx = strnlen(px, b.name_off[x + 1] - b.name_off[x]), ly = strnlen(py, b.name_off[y + 1] - b.name_off[y]);
        return lx == ly && memcmp(px, py, lx) == 0;
    };
    bool has_last = true;
    for (int i = 1; i < b.n; ++i) {
        if (has_last) {
            if (same_name(i, i - 1)) { sep[1].push_back(i - 1); sep[1].push_back(i); has_last = false; }
            else sep[0].push_back(i - 1);
        } else has_last = true;
    }
    if (has_last && b.n > 0) sep[0].push_back(b.n - 1);
    if (ma.verbose >= 3) {
        std::lock_guard<std::mutex> l(log_m);
        fprintf(log, "[M::%s] %d single-end sequences; %d paired-end sequences\n", "process", (int)sep[0].size(), (int)sep[1].size());
    }
    std::vector<std::string> sam(b.n);
    for (int k = 0; k < 2; ++k) {
        if (sep[k].empty()) continue;
        ReadBatch sub;
        sub.clear();
        FastxRecord r;
        for (int i : sep[k]) {
            r.name.assign(b.names.data() + b.name_off[i], b.name_off[i + 1] - b.name_off[i]);
            r.comment.assign(b.comments.data() + b.cmt_off[i], b.cmt_off[i + 1] - b.cmt_off[i]);
            r.seq.assign(b.bases.data() + b.seq_off[i], b.seq_off[i + 1] - b.seq_off[i]);
            if (b.has_qual[i]) r.qual.assign(b.qual.data() + b.seq_off[i], b.seq_off[i + 1] - b.seq_off[i]); else r.qual.clear();
            sub.add(r, ma.copy_comment, b.first[i], b.read_group[i], b.pattern[i]);
        }
        MemArgs mk = ma;
        if (k) mk.opt.flag |= F_PE; else mk.opt.flag &= ~F_PE;
        BatchResult rk;
        rk.want_text = true; rk.rg_id = ma.rg_id;
        const double ta = now_sec();
        aligner.align(mk.opt, sub, n_processed + (k ? (int64_t)sep[0].size() : 0), k && ma.have_pes0 ? ma.pes0 : nullptr, rk, slot);
        if (ma.verbose >= 3) {
            std::lock_guard<std::mutex> l(log_m);
            fputs(rk.log_text.c_str(), log);
            fprintf(log, "[M::%s] Processed %d reads in %.3f real sec\n", "mem_process_seqs", sub.n, now_sec() - ta);
        }
        if (rk.have_text) {
            const char *text = reinterpret_cast<const char *>(rk.text.data());
            for (int q = 0; q < sub.n; ++q) sam[sep[k][q]].assign(text + rk.text_off[q], rk.text_off[q + 1] - rk.text_off[q]);
        } else {
            const SamView v = make_sam_view(mk, idx, sub, rk);
            EntryStats st;
            for (int q = 0; q < sub.n; ++q) format_entry_append(v, k != 0, q, sam[sep[k][q]], st);
        }
        res.ms_h2d += rk.ms_h2d; res.ms_kernels += rk.ms_kernels; res.ms_d2h += rk.ms_d2h;
        for (int s = 0; s < 8; ++s) res.ms_stage[s] += rk.ms_stage[s];
        res.n_seeds += rk.n_seeds; res.h2d_bytes += rk.h2d_bytes; res.d2h_bytes += rk.d2h_bytes;
        if (k) memcpy(res.pes, rk.pes, sizeof res.pes);
    }
    size_t total = 0;
    for (const std::string &s : sam) total += s.size();
    res.text.resize_uninit(total);
    res.text_off.resize(b.n + 1);
    const SamStats zero = {0, 0, 0, 0, 0};
    res.stats.assign(b.n, zero);
    size_t o = 0;
    for (int i = 0; i < b.n; ++i) {
        res.text_off[i] = (uint32_t)o;
        memcpy(res.text.data() + o, sam[i].data(), sam[i].size());
        o += sam[i].size();
    }
    res.text_off[b.n] = (uint32_t)o;
    res.have_text = true;
}

// Three overlapped stages, like the reference's kt_pipeline (fastmap.c:352) but with the GPU in the
// middle: [read + convert-pattern bookkeeping] -> [device batch] -> [SAM text + arbiter + write].
// Each stage handles batches strictly in input order, so output order and n_processed are unchanged.
int run_mem(const MemArgs &ma, const HostIndex &idx, BatchAligner &aligner, FILE *out, FILE *log, RunSummary *summary, BamWriter *bam)
{
    double t0 = now_sec();
    const int n_dev = std::max(1, aligner.devices());
    const int parse_threads = host_parse_threads(n_dev);
    FastxReader r1(ma.fq1, parse_threads, ma.opt.undirectional != 0);
    std::unique_ptr<FastxReader> r2;
    if (!ma.fq2.empty()) r2.reset(new FastxReader(ma.fq2, parse_threads, ma.opt.undirectional != 0));
    std::string hdr = sam_header(idx, ma);
    const int shard_count = ma.shard_count > 1 ? ma.shard_count : 1, shard_index = ma.shard_index;
    size_t out_bytes = 0;
    if (bam && shard_count > 1) throw std::runtime_error("[E::run_mem] BAM output of a sharded run is written after the merge");
    // BGZF blocks straight from the device: not when the user supplied @SQ lines (the device resolves contig names through the index's
    // own table) or asked for the FASTQ comments (free text the device formatter does not carry)
    const bool dev_bam = bam && bam->device_blocks() && !ma.copy_comment && !(ma.have_hdr && ma.hdr_line.find("@SQ") != std::string::npos);
    // records leave as SAM text on `out` or, with a BamWriter, as BAM records (the `| stream_bam` half of the reference pipeline)
    auto put_out = [&](const char *p, size_t n) {
        if (bam) bam->records(p, n);
        else if (fwrite(p, 1, n, out) != n) throw std::runtime_error("[E::run_mem] writing the SAM stream failed");
    };
    if (shard_index == 0) {
        if (bam) bam->header(hdr); else put_out(hdr.data(), hdr.size());
        out_bytes += hdr.size();
    }
    FILE *parts = nullptr;
    if (shard_count > 1 && !ma.shard_parts.empty()) {
        parts = fopen(ma.shard_parts.c_str(), "w");
        if (!parts) throw std::runtime_error("[E::run_mem] cannot write " + ma.shard_parts);
        if (shard_index == 0) fprintf(parts, "-1\t0\t%zu\n", hdr.size());
    }
    RunSummary sum;
    // formatter threads: this process's share of the cores (one process per GPU under torchrun / the shard launcher),
    // minus the reader, the parser threads and the device threads, which must never wait for a core
    int host_threads = host_core_share();
    if (ma.shard_count > 1 && !getenv("LOCAL_WORLD_SIZE")) host_threads /= ma.shard_count;
    host_threads -= 5;
    if (const char *e = getenv("BSB_HOST_THREADS")) host_threads = atoi(e);
    if (host_threads < 1) host_threads = 1;
    if (host_threads > 32) host_threads = 32;
    // three batches in flight per device: one slot's host round trips and copies are covered by the other two
    int slots_per_dev = aligner.slots() / n_dev;
    if (n_dev >= 4) slots_per_dev = std::min(slots_per_dev, 2);   // measured at 8 GPUs: 24 device threads in one process cost more than the third slot hides
    if (const char *e = getenv("BSB_GPU_SLOTS")) slots_per_dev = std::max(1, std::min(slots_per_dev, atoi(e)));
    const int slot_stride = aligner.slots() / n_dev;   // slot index of device d's first context
    const int n_slots = slots_per_dev * n_dev;
    Channel<std::unique_ptr<Job>> q_plan(2), q_free(64);
    std::vector<std::unique_ptr<Channel<std::unique_ptr<Job>>>> q_read;   // one per device: batch b goes to device b mod G
    for (int d = 0; d < n_dev; ++d) q_read.emplace_back(new Channel<std::unique_ptr<Job>>((size_t)std::max(2, slots_per_dev)));
    OrderedDone q_done;
    const int n_jobs = 5 + 2 * n_slots;
    std::thread t_prefill;   // page-locks the transfer buffers of the other jobs while the first batches run
    for (int k = 0; k < n_jobs; ++k) q_free.push(std::unique_ptr<Job>(new Job)); // recycled: their buffers stay mapped and sized
    std::string fail;
    std::mutex fail_m;
    auto set_fail = [&](const std::string &w) { std::lock_guard<std::mutex> l(fail_m); if (fail.empty()) fail = w; };
    std::mutex log_m;

    // BSB_RESIDENT_BENCH (measurement): read and upload every batch first, then release them to the device slots at
    // once; sec_resident = wall time from that moment to the last batch leaving the device
    const bool resident = getenv("BSB_RESIDENT_BENCH") != nullptr;
    std::vector<std::unique_ptr<Job>> held, finished;
    double t_res0 = 0, t_res1 = 0;
    std::mutex res_m;
    double sec_plan = 0, sec_fill = 0;
    // reader, first half: cut the input into the reference's batches (serial walk over the parsers' records)
    std::thread t_read([&] {
        try {
            int64_t n_processed = 0;
            long batch_id = 0, seq = 0;
            for (;;) {
                std::unique_ptr<Job> j;
                if (resident) j.reset(new Job);
                else if (!q_free.pop(j)) break;
                double tr = now_sec();
                if (!plan_batch(ma.actual_chunk_size(), &r1, r2.get(), ma.opt.undirectional, ma.opt.substitution_proportion, j->plan)) break;
                sec_plan += now_sec() - tr;
                const int n = (int)j->plan.n_entries;
                j->n_processed = n_processed;
                j->batch_id = batch_id++;
                n_processed += n;
                j->seq = -1;
                if (j->batch_id % shard_count == shard_index) j->seq = seq++;   // else: another GPU's batch, only its blocks are released
                q_plan.push(std::move(j));
                { std::lock_guard<std::mutex> l(fail_m); if (!fail.empty()) break; }
            }
        } catch (const std::exception &e) { set_fail(e.what()); }
        q_plan.close();
    });
    // reader, second half: copy the batch into its flat (page-locked) arrays while the next one is being cut
    std::thread t_fill([&] {
        try {
            const int nt = host_fill_threads(n_dev);
            std::unique_ptr<Job> j;
            while (q_plan.pop(j)) {
                if (j->seq < 0) {
                    r1.release_until(j->plan.mark1);
                    if (r2) r2->release_until(j->plan.mark2);
                    q_free.push(std::move(j));
                    continue;
                }
                double tr = now_sec();
                fill_batch(j->plan, &r1, r2.get(), ma.copy_comment, nt, j->batch);
                sec_fill += now_sec() - tr;
                if (ma.verbose >= 3) { std::lock_guard<std::mutex> l(log_m); fprintf(log, "[M::%s] read %d sequences (%ld bp)...\n", "process", j->batch.n, (long)j->batch.n_bases); }
                const int dev = (int)(j->seq % n_dev);
                if (resident) { aligner.preload(j->batch, dev * slot_stride); held.push_back(std::move(j)); continue; }
                q_read[dev]->push(std::move(j));
            }
            if (resident) {   // every batch is in HBM: release them all at once, each device's in its own order
                if (!held.empty() && g_host_alloc.prefill) {
                    // the held batches have taken the pool's blocks: page-lock the result buffers of the batches that will be
                    // in flight now, not inside the timed region (text: about 500 bytes per entry; the rest is exact)
                    const ReadBatch &b0 = held[0]->batch;
                    const size_t n0 = (size_t)b0.n, s_reads = n0 * sizeof(ReadOut), s_off = (n0 + 1) * 4, s_stats = n0 * sizeof(SamStats), s_text = n0 * 520;
                    const size_t bytes[4] = {s_reads + s_reads / 4 + 4096, s_off + s_off / 4 + 4096, s_stats + s_stats / 4 + 4096, s_text + s_text / 4 + 4096};
                    const int want = 2 * n_slots + 2, count[4] = {want, want, want, want};
                    g_host_alloc.prefill(bytes, count, 4, true);
                }
                std::vector<std::vector<std::unique_ptr<Job>>> per_dev(n_dev);
                for (auto &h : held) { const int d = (int)(h->seq % n_dev); per_dev[d].push_back(std::move(h)); }
                held.clear();
                std::vector<std::thread> rel;
                t_res0 = now_sec();
                for (int d = 0; d < n_dev; ++d)
                    rel.emplace_back([&, d] { for (auto &h : per_dev[d]) q_read[d]->push(std::move(h)); });
                for (auto &t : rel) t.join();
            }
        } catch (const std::exception &e) {
            set_fail(e.what());
            std::unique_ptr<Job> j;
            while (q_plan.pop(j)) {}
        }
        for (auto &q : q_read) q->close();
    });
    // one host thread per device slot: batches are taken in input order and may finish out of order
    std::vector<std::thread> t_gpu;
    for (int slot = 0; slot < n_slots; ++slot) q_done.add_producer();
    for (int k = 0; k < n_slots; ++k)
        t_gpu.emplace_back([&, k] {
            const int dev = k / slots_per_dev, slot = dev * slot_stride + k % slots_per_dev;
            Channel<std::unique_ptr<Job>> &mine = *q_read[dev];
            try {
                std::unique_ptr<Job> j;
                while (mine.pop(j)) {
                    double ta = now_sec();
                    j->res.want_text = true; j->res.rg_id = ma.rg_id;   // SAM text from the device when the aligner can produce it
                    const bool smart = (ma.opt.flag & F_SMARTPE) != 0;
                    j->res.want_bam = dev_bam && !smart;               // ... or, for a BAM file, the finished BGZF blocks
                    if (smart) {
                        for (int k = 0; k < 8; ++k) j->res.ms_stage[k] = 0;
                        j->res.ms_h2d = j->res.ms_kernels = j->res.ms_d2h = 0; j->res.n_seeds = j->res.h2d_bytes = j->res.d2h_bytes = 0;
                        align_smart_pairs(ma, idx, aligner, j->batch, j->n_processed, j->res, slot, log, log_m);
                    } else
                    aligner.align(ma.opt, j->batch, j->n_processed, ma.have_pes0 ? ma.pes0 : nullptr, j->res, slot);
                    j->sec_align = now_sec() - ta;
                    if (!j->res.log_text.empty()) { std::lock_guard<std::mutex> l(log_m); fputs(j->res.log_text.c_str(), log); j->res.log_text.clear(); }
                    { std::lock_guard<std::mutex> l(res_m); t_res1 = std::max(t_res1, now_sec()); }
                    if (ma.verbose >= 3 && !smart) { std::lock_guard<std::mutex> l(log_m); fprintf(log, "[M::%s] Processed %d reads in %.3f real sec\n", "mem_process_seqs", j->batch.n, j->sec_align); }
                    q_done.push(std::move(j));
                }
            } catch (const std::exception &e) {
                set_fail(e.what());
                q_done.abort();
                std::unique_ptr<Job> j;
                while (mine.pop(j)) {} // drain so that the reader can finish
            }
            q_done.producer_done();
        });
    {
        // SAM text: each formatter thread appends the records of a run of consecutive entries to ONE buffer, so that in
        // the common case (every entry printed as formatted, in input order) the buffers are written out as they are
        struct Piece { std::string buf; std::vector<size_t> end; int lo = 0, hi = 0; };
        std::vector<Piece> pieces(host_threads);
        std::vector<EntryStats> st;
        std::vector<int> emit;
        std::vector<uint8_t> rewrite;
        std::string tmp;
        RawBuf text(false);   // SAM text never crosses PCIe: ordinary memory
        std::unique_ptr<Job> j;
        while (q_done.pop(j)) {
            try {
                const ReadBatch &batch = j->batch;
                if (sum.n_batches == 0 && g_host_alloc.prefill && !resident && !t_prefill.joinable()) {
                    const size_t s_bases = batch.bases.capacity(), s_names = batch.names.capacity(), s_reads = j->res.reads.size() * sizeof(ReadOut);
                    const size_t s_arena = j->res.have_bam ? j->res.bam.size() : j->res.have_text ? j->res.text.size() : j->res.arena.size();
                    const size_t s_off = j->res.text_off.size() * 4, s_stats = j->res.stats.size() * sizeof(SamStats);
                    t_prefill = std::thread([=] {
                        // every job in flight holds: bases, qualities, names; result text (or arena), per-read records, offsets, statistics
                        const size_t bytes[7] = {s_bases, s_bases, s_names, s_arena + s_arena / 4 + 4096, s_reads + s_reads / 4 + 4096,
                                                 s_off + s_off / 4 + 4096, s_stats + s_stats / 4 + 4096};
                        const int count[7] = {n_jobs, n_jobs, n_jobs, n_jobs, n_jobs, s_off ? n_jobs : 0, s_off ? n_jobs : 0};
                        g_host_alloc.prefill(bytes, count, 7, false);
                    });
                }
                sum.sec_align += j->sec_align;
                sum.add_timing(j->res);
                double tf = now_sec();
                st.resize(batch.n);
                MapStats ms;
                size_t total = 0;
                double tw;
                if (j->res.have_bam) {
                    // records, arbiter and compression ran on the device: the blocks go to the file as they are
                    const BatchResult &R = j->res;
                    ms.reads = (long)R.bam_counts[0]; ms.alignments = (long)R.bam_counts[1]; ms.wc2t = (long)R.bam_counts[2]; ms.wg2a = (long)R.bam_counts[3];
                    ms.cc2t = (long)R.bam_counts[4]; ms.cg2a = (long)R.bam_counts[5]; ms.unaligned = (long)R.bam_counts[6]; ms.bs_ambiguous = (long)R.bam_counts[7];
                    tw = now_sec();
                    total = R.bam.size();
                    bam->blocks(R.bam.data(), total, R.bam_raw_bytes, R.bam_records);
                } else
                if (j->res.have_text) {
                    // the records were formatted on the device: only the arbiter runs here
                    const BatchResult &R = j->res;
                    parallel_for(host_threads, batch.n, [&](int i) {
                        const SamStats &x = R.stats[i];
                        st[i].alignment_score = x.alignment_score; st[i].mapped = x.mapped; st[i].bs_conflict = x.bs_conflict; st[i].crick = x.crick; st[i].paired = x.paired;
                    });
                    sam_sort_plan_parallel(batch, st, emit, rewrite, ms, std::min(host_threads, 8));
                    bool as_is = (int)emit.size() == batch.n;
                    for (int k = 0; as_is && k < batch.n; ++k) as_is = emit[k] == k && !rewrite[k];
                    const char *text = reinterpret_cast<const char *>(R.text.data());
                    if (as_is) {
                        tw = now_sec();
                        total = R.text.size();
                        put_out(text, total);
                    } else {
                        // entries were dropped or rewritten (undirectional libraries): the text leaves in runs of consecutive
                        // entries straight from the device's buffer -- nothing is assembled, only the rewritten records are made here
                        tw = now_sec();
                        size_t k = 0;
                        const size_t m = emit.size();
                        while (k < m) {
                            const int i = emit[k];
                            if (rewrite[i]) { set_unmapped(batch, i, st[i], tmp); put_out(tmp.data(), tmp.size()); total += tmp.size(); ++k; continue; }
                            size_t k2 = k;
                            while (k2 + 1 < m && emit[k2 + 1] == emit[k2] + 1 && !rewrite[emit[k2 + 1]]) ++k2;
                            const size_t len = R.text_off[emit[k2] + 1] - R.text_off[i];
                            put_out(text + R.text_off[i], len);
                            total += len;
                            k = k2 + 1;
                        }
                    }
                } else {
                const int nt = std::max(1, std::min(host_threads, batch.n / 256 + 1));
                const int chunk = (batch.n + nt - 1) / nt;
                for (int t = 0; t < nt; ++t) { pieces[t].lo = std::min(batch.n, t * chunk); pieces[t].hi = std::min(batch.n, (t + 1) * chunk); }
                const SamView view = make_sam_view(ma, idx, batch, j->res);
                const bool is_pe = (ma.opt.flag & F_PE) != 0;
                parallel_for(nt, nt, [&](int t) {
                    Piece &pc = pieces[t];
                    pc.buf.clear(); pc.end.resize(pc.hi - pc.lo);
                    for (int i = pc.lo; i < pc.hi; ++i) {
                        format_entry_append(view, is_pe, i, pc.buf, st[i]);
                        pc.end[i - pc.lo] = pc.buf.size();
                    }
                }, 1);
                sam_sort_plan_parallel(batch, st, emit, rewrite, ms, std::min(host_threads, 8));
                bool as_is = (int)emit.size() == batch.n;
                for (int k = 0; as_is && k < batch.n; ++k) as_is = emit[k] == k && !rewrite[k];
                if (as_is) {
                    tw = now_sec();
                    for (int t = 0; t < nt; ++t) { put_out(pieces[t].buf.data(), pieces[t].buf.size()); total += pieces[t].buf.size(); }
                } else {   // some entries dropped or rewritten by the arbiter: assemble the text entry by entry
                    tmp.clear();
                    auto span = [&](int i, const char *&p, size_t &l) {
                        const Piece &pc = pieces[i / chunk];
                        const size_t b0 = i == pc.lo ? 0 : pc.end[i - pc.lo - 1];
                        p = pc.buf.data() + b0; l = pc.end[i - pc.lo] - b0;
                    };
                    std::string all;
                    for (int i : emit) {
                        if (rewrite[i]) { set_unmapped(batch, i, st[i], tmp); all += tmp; }
                        else { const char *p; size_t l; span(i, p, l); all.append(p, l); }
                    }
                    total = all.size();
                    tw = now_sec();
                    put_out(all.data(), total);
                }
                }
                if (parts) fprintf(parts, "%ld\t%zu\t%zu\n", j->batch_id, out_bytes, total);
                out_bytes += total;
                sum.sec_format += tw - tf; sum.sec_write += now_sec() - tw;
                {
                    std::lock_guard<std::mutex> l(log_m);
                    fprintf(log, "BSStat TotalReads: %ld\n", ms.reads);
                    fprintf(log, "BSStat TotalAlignments: %ld\n", ms.alignments);
                    fprintf(log, "BSStat W_C2T: %ld\n", ms.wc2t);
                    fprintf(log, "BSStat W_G2A: %ld\n", ms.wg2a);
                    fprintf(log, "BSStat C_C2T: %ld\n", ms.cc2t);
                    fprintf(log, "BSStat C_G2A: %ld\n", ms.cg2a);
                    fprintf(log, "BSStat Unaligned: %ld\n", ms.unaligned);
                    fprintf(log, "BSStat BSAmbiguous: %ld\n", ms.bs_ambiguous);
                }
                sum.stats.add(ms);
                ++sum.n_batches;
                sum.n_entries += batch.n;
            } catch (const std::exception &e) { set_fail(e.what()); }
            if (resident) {   // the page-locked host buffers go back to the pool now, for the batches still to come
                j->res.arena.reset(); j->res.reads.reset(); j->res.text.reset(); j->res.bam.reset(); j->res.text_off.reset(); j->res.stats.reset(); j->batch.bases.reset();
                finished.push_back(std::move(j));
            }   // (device inputs are freed after the run: cudaFree synchronises the device)
            else q_free.push(std::move(j));
        }
    }
    q_free.close();
    for (auto &t : t_gpu) t.join();
    t_fill.join();
    t_read.join();
    if (t_prefill.joinable()) t_prefill.join();
    for (auto &h : finished) aligner.unload(h->batch);
    finished.clear();
    if (out) fflush(out);
    if (parts) fclose(parts);
    sum.sec_read = std::max(sec_plan, sec_fill);   // the slower of the reader's two overlapped halves
    sum.sec_plan = sec_plan; sum.sec_fill = sec_fill;
    sum.sec_resident = resident ? t_res1 - t_res0 : 0;
    sum.sec_total = now_sec() - t0;
    if (summary) *summary = sum;
    if (!fail.empty()) throw std::runtime_error(fail);
    return 0;
}

} // namespace bsb
