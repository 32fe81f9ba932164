// bsb_index_build.cu -- GPU construction of the index files `bsbolt Index` produces with `bwa index`
// (SURVEY 8 f-1): byte-identical .pac .opac .ann .amb .bwt .sa for references below 2^31 bases
// (text of up to 2^32-2 symbols; the human-scale 12.4 G-symbol text still needs the offline CPU build).
//
//   pack_fasta()        <- bns_fasta2bntseq + add1            (bntseq.c:239-361)   host
//   build_sa_bwt()      <- bwt_bwtgen2 / bwt_pac2bwt          (bwt_gen.c, bwtindex.c:64-118): the BWT of
//                          P.revcomp(P) is unique, so instead of BWT-SW's incremental merge the suffix
//                          array is built on the device by prefix doubling over CUB radix sorts
//   interleave + sample <- bwt_bwtupdate_core, bwt_cal_sa(32) (bwtindex.c:151-173, bwt.c:61-84)
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <stdexcept>
#include <string>
#include <vector>
#include "bsb_index_build.h"

namespace bsb {

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("[E::bsb_index_build] CUDA error: ") + cudaGetErrorString(e_) + " (" #x ") at line " + std::to_string(__LINE__)); } while (0)

// ------------------------------------------------------------------------------------------------
// host: FASTA -> 2-bit pac / opac, .ann, .amb
// ------------------------------------------------------------------------------------------------
static const unsigned char kNt4[256] = {
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};

struct FaContig { std::string name, comment, seq; };

static void read_fasta(const std::string &path, std::vector<FaContig> &out)
{
    gzFile fp = gzopen(path.c_str(), "r");
    if (!fp) throw std::runtime_error("[E::bsb_index_build] fail to open file '" + path + "'");
    gzbuffer(fp, 1 << 20);
    std::vector<char> buf(1 << 20);
    std::string line;
    bool in_seq = false;
    auto handle = [&](const std::string &l) {
        if (!l.empty() && l[0] == '>') {
            out.emplace_back();
            size_t p = 1;
            while (p < l.size() && !isspace((unsigned char)l[p])) ++p;
            out.back().name = l.substr(1, p - 1);
            if (p < l.size()) out.back().comment = l.substr(p + 1);
            in_seq = true;
        } else if (in_seq) out.back().seq += l;
    };
    for (;;) {
        int n = gzread(fp, buf.data(), (unsigned)buf.size());
        if (n <= 0) break;
        int s = 0;
        for (int i = 0; i < n; ++i)
            if (buf[i] == '\n') {
                line.append(buf.data() + s, i - s);
                if (line.size() > 1 && line.back() == '\r') line.pop_back();
                handle(line);
                line.clear();
                s = i + 1;
            }
        line.append(buf.data() + s, n - s);
    }
    if (!line.empty()) handle(line);
    gzclose(fp);
}

struct Hole { int64_t offset; int32_t len; char amb; };

// one pass of bns_fasta2bntseq(fa, prefix, for_only=1, bs): forward strands, watson contigs then crick contigs
static void pack_pass(const std::vector<FaContig> &ctg, int conversion, std::vector<uint8_t> &pac, int64_t &l_pac,
                      std::vector<Hole> &holes, std::vector<int> &n_ambs)
{
    srand48(11);
    int64_t total = 0;
    for (auto &c : ctg) total += (int64_t)c.seq.size();
    total *= 2;
    pac.assign((size_t)(total / 4 + 2), 0);
    holes.clear(); n_ambs.assign(ctg.size() * 2, 0);
    l_pac = 0;
    Hole *q = nullptr;
    for (int is_crick = 0; is_crick < 2; ++is_crick) {
        for (size_t ci = 0; ci < ctg.size(); ++ci) {
            const std::string &s = ctg[ci].seq;
            const int64_t off = l_pac;
            int lasts = 0;
            for (size_t i = 0; i < s.size(); ++i) {
                int c = kNt4[(unsigned char)s[i]];
                if (conversion) {
                    if (is_crick) { if (c == 2) c = 0; }
                    else { if (c == 1) c = 3; }
                }
                if (c >= 4) {
                    if (lasts == s[i] && q) ++q->len;
                    else {
                        holes.push_back(Hole{off + (int64_t)i, 1, s[i]});
                        q = &holes.back();
                        ++n_ambs[is_crick * ctg.size() + ci];
                    }
                    c = (int)(lrand48() & 3);
                }
                lasts = s[i];
                pac[(size_t)(l_pac >> 2)] |= (uint8_t)(c << ((~l_pac & 3) << 1));
                ++l_pac;
            }
        }
    }
}

static void write_pac(const std::string &fn, const std::vector<uint8_t> &pac, int64_t l_pac)
{
    FILE *fp = fopen(fn.c_str(), "wb");
    if (!fp) throw std::runtime_error("[E::bsb_index_build] cannot write " + fn);
    fwrite(pac.data(), 1, (size_t)((l_pac >> 2) + ((l_pac & 3) == 0 ? 0 : 1)), fp);
    unsigned char ct;
    if (l_pac % 4 == 0) { ct = 0; fwrite(&ct, 1, 1, fp); }
    ct = (unsigned char)(l_pac % 4);
    fwrite(&ct, 1, 1, fp);
    fclose(fp);
}

static void write_ann_amb(const std::string &prefix, const std::vector<FaContig> &ctg, int64_t l_pac,
                          const std::vector<Hole> &holes, const std::vector<int> &n_ambs)
{
    FILE *fp = fopen((prefix + ".ann").c_str(), "w");
    if (!fp) throw std::runtime_error("[E::bsb_index_build] cannot write " + prefix + ".ann");
    fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)ctg.size() * 2, 11u);
    int64_t off = 0;
    for (int is_crick = 0; is_crick < 2; ++is_crick)
        for (size_t ci = 0; ci < ctg.size(); ++ci) {
            const FaContig &c = ctg[ci];
            fprintf(fp, "%d %s%s", 0, c.name.c_str(), is_crick ? "_crick_bs" : "");
            fprintf(fp, " %s\n", c.comment.empty() ? "(null)" : c.comment.c_str());
            fprintf(fp, "%lld %d %d\n", (long long)off, (int)c.seq.size(), n_ambs[is_crick * ctg.size() + ci]);
            off += (int64_t)c.seq.size();
        }
    fclose(fp);
    fp = fopen((prefix + ".amb").c_str(), "w");
    if (!fp) throw std::runtime_error("[E::bsb_index_build] cannot write " + prefix + ".amb");
    fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)ctg.size() * 2, (unsigned)holes.size());
    for (const Hole &h : holes) fprintf(fp, "%lld %d %c\n", (long long)h.offset, h.len, h.amb);
    fclose(fp);
}

// ------------------------------------------------------------------------------------------------
// device: suffix array by prefix doubling
// ------------------------------------------------------------------------------------------------
__global__ void k_unpack_text(const uint8_t *pac, int64_t l_pac, uint8_t *T)
{   // T = P . revcomp(P), one byte per symbol
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * l_pac) return;
    int64_t j = i < l_pac ? i : 2 * l_pac - 1 - i;
    int c = pac[j >> 2] >> ((~j & 3) << 1) & 3;
    T[i] = (uint8_t)(i < l_pac ? c : 3 - c);
}

// key of the first 21 symbols (3 bits each: symbol+1, 0 past the end) for positions 0..n
__global__ void k_init_keys(const uint8_t *T, uint32_t n, uint64_t *key, uint32_t *val)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint64_t k = 0;
#pragma unroll
    for (int d = 0; d < 21; ++d) {
        uint64_t p = (uint64_t)i + d;
        uint64_t s = p < n ? (uint64_t)T[p] + 1 : 0;
        k = k << 3 | s;
    }
    key[i] = k; val[i] = i;
}

__global__ void k_flag_heads(const uint64_t *key, uint32_t m, uint32_t *head)
{   // head[j] = j if key[j] starts a new group else 0
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    head[j] = (j == 0 || key[j] != key[j - 1]) ? j : 0;
}

__global__ void k_scatter_rank(const uint32_t *sa, const uint32_t *grp, uint32_t m, uint32_t *rank)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) rank[sa[j]] = grp[j];
}

__global__ void k_count_heads(const uint64_t *key, uint32_t m, unsigned long long *cnt)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    int f = (j < m && (j == 0 || key[j] != key[j - 1])) ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt, (unsigned long long)__popc(b));
}

__global__ void k_double_keys(const uint32_t *sa, const uint32_t *rank, uint32_t m, uint32_t h, uint64_t *key)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    uint32_t i = sa[j];
    uint64_t p = (uint64_t)i + h;
    uint32_t r2 = p < m ? rank[p] + 1 : 0;   // suffixes shorter than h are already unique
    key[j] = (uint64_t)rank[i] << 32 | r2;
}

struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

// stored BWT symbol j (the '$' row removed) and per-128-block symbol counts
__global__ void k_bwt_syms(const uint8_t *T, const uint32_t *sa, uint32_t n, uint32_t primary, uint8_t *bw)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t r = j + (j >= primary ? 1u : 0u);
    uint32_t p = sa[r];            // p > 0 here: the row with p == 0 is `primary`
    bw[j] = T[p - 1];
}

__global__ void k_block_counts(const uint8_t *bw, uint32_t n, uint32_t n_blk, uint32_t *c0, uint32_t *c1, uint32_t *c2, uint32_t *c3)
{   // one warp per block of 128 symbols
    uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blk >= n_blk) return;
    uint32_t a = 0, c = 0, g = 0, t = 0;
    for (int k = 0; k < 4; ++k) {
        uint64_t j = (uint64_t)blk * 128 + lane * 4 + k;
        if (j < n) { int s = bw[j]; a += s == 0; c += s == 1; g += s == 2; t += s == 3; }
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o);
        g += __shfl_xor_sync(0xffffffffu, g, o); t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane == 0) { c0[blk] = a; c1[blk] = c; c2[blk] = g; c3[blk] = t; }
}

// interleaved layout: per 128 symbols 4 x u64 cumulative counts + 8 x u32 packed symbols; one extra count block at the end
__global__ void k_interleave(const uint8_t *bw, uint32_t n, uint32_t n_blk, const uint64_t *s0, const uint64_t *s1, const uint64_t *s2,
                             const uint64_t *s3, uint32_t *out, uint64_t out_words)
{
    uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk > n_blk) return;
    // words before block blk: blk*16 (full blocks) -- the final count block follows the last (possibly partial) block
    uint64_t base;
    if (blk < n_blk) base = (uint64_t)blk * 16;
    else base = (uint64_t)(n_blk - 1) * 16 + 8 + (((uint64_t)n - (uint64_t)(n_blk - 1) * 128 + 15) >> 4);
    if (n_blk == 0) base = 0;
    uint64_t c[4] = {s0[blk], s1[blk], s2[blk], s3[blk]};
    for (int k = 0; k < 4; ++k) { out[base + 2 * k] = (uint32_t)c[k]; out[base + 2 * k + 1] = (uint32_t)(c[k] >> 32); }
    if (blk == n_blk) return;
    uint64_t first = (uint64_t)blk * 128;
    for (int w = 0; w < 8; ++w) {
        uint64_t j0 = first + (uint64_t)w * 16;
        if (j0 >= n) break;
        uint32_t word = 0;
        for (int k = 0; k < 16; ++k) {
            uint64_t j = j0 + k;
            if (j < n) word |= (uint32_t)bw[j] << ((~k & 15) << 1);
        }
        out[base + 8 + w] = word;
    }
}

__global__ void k_sample_sa(const uint32_t *sa, uint32_t m, int intv, uint64_t *out)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t r = (uint64_t)j * intv;
    if (r < m) out[j] = sa[r];
}

template <class T> struct DBuf {
    T *p = nullptr;
    void alloc(size_t n) { CKB(cudaMalloc((void **)&p, (n + 8) * sizeof(T))); }
    ~DBuf() { if (p) cudaFree(p); }
};

static inline unsigned cdivu(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

void index_build(const std::string &fasta, const std::string &prefix, int device, IndexBuildStats *stats)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("[E::bsb_index_build] no CUDA device available; the index builder has no CPU fallback");
    CKB(cudaSetDevice(device));
    std::vector<FaContig> ctg;
    read_fasta(fasta, ctg);
    if (ctg.empty()) throw std::runtime_error("[E::bsb_index_build] no sequences in " + fasta);
    std::vector<uint8_t> pac, opac;
    std::vector<Hole> holes, holes2;
    std::vector<int> n_ambs, n_ambs2;
    int64_t l_pac = 0, l_pac2 = 0;
    pack_pass(ctg, 1, pac, l_pac, holes, n_ambs);
    pack_pass(ctg, 0, opac, l_pac2, holes2, n_ambs2);
    if ((uint64_t)l_pac * 2 + 2 >= (1ull << 32)) throw std::runtime_error("[E::bsb_index_build] reference too long for the 32-bit device builder (needs < 2^31 bases incl. both conversions)");
    write_pac(prefix + ".pac", pac, l_pac);
    write_pac(prefix + ".opac", opac, l_pac);
    write_ann_amb(prefix, ctg, l_pac, holes2, n_ambs2);

    const uint32_t n = (uint32_t)(2 * l_pac), m = n + 1; // m suffixes including the empty one
    cudaEvent_t e0, e1;
    CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
    CKB(cudaEventRecord(e0));
    DBuf<uint8_t> d_pac, d_T, d_bw;
    d_pac.alloc(pac.size()); CKB(cudaMemcpy(d_pac.p, pac.data(), pac.size(), cudaMemcpyHostToDevice));
    d_T.alloc(n);
    k_unpack_text<<<cdivu(n, 256), 256>>>(d_pac.p, l_pac, d_T.p);
    DBuf<uint64_t> d_key, d_key2;
    DBuf<uint32_t> d_val, d_val2, d_rank, d_grp;
    d_key.alloc(m); d_key2.alloc(m); d_val.alloc(m); d_val2.alloc(m); d_rank.alloc(m); d_grp.alloc(m);
    DBuf<unsigned long long> d_cnt; d_cnt.alloc(1);
    k_init_keys<<<cdivu(m, 256), 256>>>(d_T.p, n, d_key.p, d_val.p);
    CKB(cudaGetLastError());
    cub::DoubleBuffer<uint64_t> kb(d_key.p, d_key2.p);
    cub::DoubleBuffer<uint32_t> vb(d_val.p, d_val2.p);
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int64_t)m, 0, 64);
    cub::DeviceScan::InclusiveScan(nullptr, tmp2, d_grp.p, d_grp.p, MaxOp(), (int64_t)m);
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    DBuf<uint8_t> d_tmp; d_tmp.alloc(tmp_bytes + 256);
    int rounds = 0;
    uint32_t h = 21;
    CKB(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, kb, vb, (int64_t)m, 0, 63));
    for (;;) {
        ++rounds;
        // group ranks from the sorted keys
        k_flag_heads<<<cdivu(m, 256), 256>>>(kb.Current(), m, d_grp.p);
        size_t tb = tmp_bytes;
        CKB(cub::DeviceScan::InclusiveScan(d_tmp.p, tb, d_grp.p, d_grp.p, MaxOp(), (int64_t)m));
        k_scatter_rank<<<cdivu(m, 256), 256>>>(vb.Current(), d_grp.p, m, d_rank.p);
        CKB(cudaMemset(d_cnt.p, 0, 8));
        k_count_heads<<<cdivu(m, 256), 256>>>(kb.Current(), m, d_cnt.p);
        unsigned long long groups = 0;
        CKB(cudaMemcpy(&groups, d_cnt.p, 8, cudaMemcpyDeviceToHost));
        if (groups == m) break;
        if (h >= m) throw std::runtime_error("[E::bsb_index_build] prefix doubling did not converge");
        k_double_keys<<<cdivu(m, 256), 256>>>(vb.Current(), d_rank.p, m, h, kb.Current());
        CKB(cudaGetLastError());
        tb = tmp_bytes;
        CKB(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, kb, vb, (int64_t)m, 0, 64));
        h <<= 1;
    }
    const uint32_t *d_sa = vb.Current(); // d_sa[r] = start of the r-th smallest suffix, d_sa[0] == n
    uint32_t primary = 0;
    CKB(cudaMemcpy(&primary, d_rank.p, 4, cudaMemcpyDeviceToHost)); // rank of suffix 0
    // BWT symbols, counts, interleaved layout
    d_bw.alloc(n);
    k_bwt_syms<<<cdivu(n, 256), 256>>>(d_T.p, d_sa, n, primary, d_bw.p);
    const uint32_t n_blk = (uint32_t)(((uint64_t)n + 127) / 128);
    DBuf<uint32_t> c32[4]; DBuf<uint64_t> c64[4];
    for (int k = 0; k < 4; ++k) { c32[k].alloc(n_blk + 1); c64[k].alloc(n_blk + 1); CKB(cudaMemset(c32[k].p, 0, ((size_t)n_blk + 1) * 4)); }
    k_block_counts<<<cdivu((uint64_t)n_blk * 32, 256), 256>>>(d_bw.p, n, n_blk, c32[0].p, c32[1].p, c32[2].p, c32[3].p);
    for (int k = 0; k < 4; ++k) {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, c32[k].p, c64[k].p, (int64_t)n_blk + 1);
        if (tb > tmp_bytes) throw std::runtime_error("[E::bsb_index_build] scan scratch too small");
        CKB(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, c32[k].p, c64[k].p, (int64_t)n_blk + 1));
    }
    const uint64_t bwt_words = (((uint64_t)n + 15) >> 4) + ((uint64_t)n_blk + 1) * 8;
    DBuf<uint32_t> d_out; d_out.alloc(bwt_words);
    CKB(cudaMemset(d_out.p, 0, bwt_words * 4));
    k_interleave<<<cdivu((uint64_t)n_blk + 1, 128), 128>>>(d_bw.p, n, n_blk, c64[0].p, c64[1].p, c64[2].p, c64[3].p, d_out.p, bwt_words);
    CKB(cudaGetLastError());
    std::vector<uint32_t> h_bwt(bwt_words);
    CKB(cudaMemcpy(h_bwt.data(), d_out.p, bwt_words * 4, cudaMemcpyDeviceToHost));
    uint64_t L2[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) { uint64_t tot; CKB(cudaMemcpy(&tot, c64[k].p + n_blk, 8, cudaMemcpyDeviceToHost)); L2[k + 1] = L2[k] + tot; }
    // SA samples
    const int sa_intv = 32;
    const uint64_t n_sa = ((uint64_t)n + sa_intv) / sa_intv;
    DBuf<uint64_t> d_ss; d_ss.alloc(n_sa);
    k_sample_sa<<<cdivu(n_sa, 256), 256>>>(d_sa, m, sa_intv, d_ss.p);
    std::vector<uint64_t> h_sa(n_sa);
    CKB(cudaMemcpy(h_sa.data(), d_ss.p, n_sa * 8, cudaMemcpyDeviceToHost));
    CKB(cudaEventRecord(e1));
    CKB(cudaEventSynchronize(e1));
    float ms = 0;
    CKB(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    { // .bwt (bwt_dump_bwt, bwt.c:385-394)
        FILE *fp = fopen((prefix + ".bwt").c_str(), "wb");
        if (!fp) throw std::runtime_error("[E::bsb_index_build] cannot write " + prefix + ".bwt");
        uint64_t p64 = primary;
        fwrite(&p64, 8, 1, fp); fwrite(L2 + 1, 8, 4, fp);
        fwrite(h_bwt.data(), 4, bwt_words, fp);
        fclose(fp);
    }
    { // .sa (bwt_dump_sa, bwt.c:396-407)
        FILE *fp = fopen((prefix + ".sa").c_str(), "wb");
        if (!fp) throw std::runtime_error("[E::bsb_index_build] cannot write " + prefix + ".sa");
        uint64_t p64 = primary, intv = sa_intv, sl = n;
        fwrite(&p64, 8, 1, fp); fwrite(L2 + 1, 8, 4, fp); fwrite(&intv, 8, 1, fp); fwrite(&sl, 8, 1, fp);
        fwrite(h_sa.data() + 1, 8, n_sa - 1, fp);
        fclose(fp);
    }
    if (stats) { stats->l_pac = l_pac; stats->seq_len = n; stats->rounds = rounds; stats->ms_device = ms; stats->n_contigs = (int)ctg.size(); }
}

} // namespace bsb
