// bam_emulation.h -- the device's BAM stage (bsb_bam.h + bsb_deflate.h) run phase by phase on the CPU. TEST INFRASTRUCTURE:
// the product runs these functions in k_bam_arbiter / k_bam_count / k_bam_write / k_bgzf_deflate (bsb_cuda.cu) and has no CPU
// path for them.
//
// XHost executes a block-wide phase as a loop over the "threads"; BSB_PAR_ORDER=reverse|shuffle changes the order in which
// they run, which must not change a single output byte (the kernels' phases are free of ordering assumptions).
#pragma once
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../bsbolt_b200/csrc/bsb_bam.h"
#include "../../bsbolt_b200/csrc/bsb_deflate.h"

namespace bsb {

struct XHost {
    int order = 0;          // 0 ascending, 1 descending, 2 shuffled
    unsigned seed = 12345;
    std::vector<int> perm;
    XHost()
    {
        const char *e = getenv("BSB_PAR_ORDER");
        if (e && std::string(e) == "reverse") order = 1;
        if (e && std::string(e) == "shuffle") order = 2;
    }
    template <class F> void par(int n, F f)
    {
        if (order == 0) { for (int i = 0; i < n; ++i) f(i); return; }
        if (order == 1) { for (int i = n - 1; i >= 0; --i) f(i); return; }
        perm.resize(n);
        std::iota(perm.begin(), perm.end(), 0);
        for (int i = n - 1; i > 0; --i) { seed = seed * 1103515245u + 12345u; std::swap(perm[i], perm[(seed >> 8) % (unsigned)(i + 1)]); }
        for (int i = 0; i < n; ++i) f(perm[i]);
    }
    template <class F> void par1(F f) { par(DF_CH, f); }
    template <class F> void wpar1(F f) { par(DF_CH, f); }
    void sync() {}
    void tick(int) {}
    void atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
    void atomic_xor(uint32_t *p, uint32_t v) { *p ^= v; }
    void atomic_add(uint32_t *p, uint32_t v) { *p += v; }
    void atomic_max16(uint32_t *w, uint32_t idx, uint32_t v)   // 16-bit entry idx of a word array
    {
        const int sh = (int)(idx & 1) << 4;
        if (v > (w[idx >> 1] >> sh & 0xffffu)) w[idx >> 1] = (w[idx >> 1] & ~(0xffffu << sh)) | v << sh;
    }
};

// raw bytes -> BGZF blocks. cuts: block k = raw[cuts[k], cuts[k + 1]); empty blocks are skipped (k_bgzf_deflate does the same)
inline void bgzf_emulated(const uint8_t *raw, const std::vector<uint64_t> &cuts, std::vector<uint8_t> &out)
{
    XHost x;
    static DeflateShared S;
    std::vector<uint32_t> tok(BGZF_MAX_IN + 8);
    std::vector<uint8_t> in(BGZF_MAX_IN + 16), slot(BGZF_SLOT + 16);
    for (size_t k = 0; k + 1 < cuts.size(); ++k) {
        const int len = (int)(cuts[k + 1] - cuts[k]);
        if (len < 0 || len > BGZF_MAX_IN) throw std::runtime_error("block cut of " + std::to_string(len) + " bytes");
        if (!len) continue;
        std::fill(in.begin(), in.end(), 0xa5);                     // (the bytes behind the input may be read, never used)
        const int skew = (int)(cuts[k] & 3);                       // the block starts at any byte address on the device
        memcpy(in.data() + skew, raw + cuts[k], (size_t)len);
        uint8_t *o = slot.data() + ((16 - (reinterpret_cast<uintptr_t>(slot.data()) & 15)) & 15);
        const uint32_t total = bgzf_block(x, S, in.data() + skew, len, o, tok.data());
        out.insert(out.end(), o, o + total);
    }
}
// fixed cuts every BGZF_MAX_IN bytes
inline void bgzf_emulated(const uint8_t *raw, size_t n, std::vector<uint8_t> &out)
{
    std::vector<uint64_t> cuts;
    for (size_t o = 0; o < n; o += BGZF_MAX_IN) cuts.push_back(o);
    cuts.push_back(n);
    bgzf_emulated(raw, cuts, out);
}
// cuts at entry starts (bam_block_cut): off[0..n] are the entries' offsets
inline std::vector<uint64_t> entry_cuts(const std::vector<uint32_t> &off)
{
    const int n = (int)off.size() - 1;
    uint32_t mx = 0;
    for (int i = 0; i < n; ++i) mx = std::max(mx, off[i + 1] - off[i]);
    const uint32_t q = bam_block_quantum(mx);
    const size_t nblk = ((size_t)off[n] + q - 1) / q + 1;
    std::vector<uint64_t> cuts(nblk + 1);
    for (size_t k = 0; k <= nblk; ++k)
        cuts[k] = mx <= 0xff00u / 2 ? bam_block_cut(off.data(), n, (uint64_t)k * q) : std::min<uint64_t>((uint64_t)k * q, off[n]);
    return cuts;
}

// One batch: arbiter, record sizes, records, blocks. Returns the BGZF bytes; counters and raw size through the references.
struct BamBatchIn {
    const char *names; const uint32_t *name_off; const uint8_t *first, *read_group;
    const char *bases, *qual; const uint32_t *seq_off; const uint8_t *has_qual;
    const char *text; const uint32_t *text_off; const SamStats *stats; int n;
    const char *ctg_text; const uint32_t *ctg_name_off; const int32_t *ctg_sorted; int n_ctg;
};

inline void bam_batch_emulated(const BamBatchIn &b, std::vector<uint8_t> &bgzf, MapCounters &ctr, uint64_t &raw_bytes, uint64_t &n_records)
{
    std::vector<uint8_t> code(b.n, BAM_DROP);
    BamView v;
    v.a.names = b.names; v.a.name_off = b.name_off; v.a.first = b.first; v.a.read_group = b.read_group; v.a.stats = b.stats; v.a.n = b.n;
    v.text = b.text; v.text_off = b.text_off; v.code = code.data();
    v.ctg.text = b.ctg_text; v.ctg.name_off = b.ctg_name_off; v.ctg.sorted = b.ctg_sorted; v.ctg.n = b.n_ctg;
    v.bases = b.bases; v.qual = b.qual; v.seq_off = b.seq_off; v.has_qual = b.has_qual;
    memset(&ctr, 0, sizeof ctr);
    for (int i = 0; i < b.n; ++i)                                  // k_bam_arbiter: one thread per entry, the group heads work
        if (v.a.is_head(i)) bam_arbitrate(v.a, i, code.data(), ctr);
    std::vector<uint32_t> off(b.n + 1, 0);
    for (int i = 0; i < b.n; ++i) {                                // k_bam_count
        SamCount c;
        const int rc = bam_entry(c, v, i, false);
        if (rc) throw std::runtime_error(std::string("[E::bam_encode] ") + bam_strerror(rc) + " (entry " + std::to_string(i) + ")");
        off[i + 1] = off[i] + (uint32_t)c.n;
    }
    std::vector<uint8_t> raw(off[b.n] + 16);
    n_records = 0;
    for (int i = 0; i < b.n; ++i) {                                // k_bam_write
        SamWrite w = {reinterpret_cast<char *>(raw.data() + off[i])};
        int nr = 0;
        const int rc = bam_entry(w, v, i, true, &nr);
        if (rc || (uint64_t)(w.p - reinterpret_cast<char *>(raw.data())) != off[i + 1]) throw std::runtime_error("[E::bam_encode] sizing and writing disagree");
        n_records += (uint64_t)nr;
    }
    raw_bytes = off[b.n];
    bgzf_emulated(raw.data(), entry_cuts(off), bgzf);             // k_bgzf_deflate + gather
}

} // namespace bsb
