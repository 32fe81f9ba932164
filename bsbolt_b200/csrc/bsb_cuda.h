// bsb_cuda.h -- the GPU batch aligner (implemented in bsb_cuda.cu; CUDA only, no CPU fallback)
#pragma once
#include <vector>
#include "host_mem.h"

namespace bsb {

class CudaAligner : public BatchAligner {
public:
    // uploads the index to HBM of `device`; throws if no CUDA device is usable
    CudaAligner(const HostIndex &idx, int device);
    ~CudaAligner() override;
    static constexpr int kSlots = 3;   // batches that can be in flight on the device at once (one host thread each)
    int slots() const override { return kSlots; }
    void align(const Opt &opt, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int slot = 0) override;
    void preload(ReadBatch &b, int slot = 0) override;
    void unload(ReadBatch &b) override;
    long kernel_launches() const;   // kernels launched by this object so far
    int device() const;
    size_t index_bytes() const;     // HBM held by the resident index
    int verbose = 3;
private:
    struct Impl;
    Impl *im_;
};

// random 32-byte sector reads over a buffer of about footprint_bytes: GB/s with independent loads and with one dependent load per thread
void random_sector_peak(int device, size_t footprint_bytes, double *gbs_independent, double *gbs_chase);

// Several GPUs behind one BatchAligner: one CudaAligner (one resident copy of the index) per device, kSlots batch
// contexts each. Nothing is exchanged between the devices -- reads are independent given the index, the options, the
// per-batch insert-size statistics and the global read index -- so there is no collective; the pipeline (run_mem) cuts the
// input ONCE, deals the batches round-robin and writes the results in input order.
class MultiAligner : public BatchAligner {
public:
    explicit MultiAligner(std::vector<CudaAligner *> per_device) : dev_(std::move(per_device)) {}
    int devices() const override { return (int)dev_.size(); }
    int slots() const override { return (int)dev_.size() * CudaAligner::kSlots; }
    void align(const Opt &opt, const ReadBatch &b, int64_t n_processed, const PeStat *pes0, BatchResult &out, int slot = 0) override
    {
        dev_[slot / CudaAligner::kSlots]->align(opt, b, n_processed, pes0, out, slot % CudaAligner::kSlots);
    }
    void preload(ReadBatch &b, int slot = 0) override { owner_of_input(b) = dev_[slot / CudaAligner::kSlots]; dev_[slot / CudaAligner::kSlots]->preload(b, 0); }
    void unload(ReadBatch &b) override { if (CudaAligner *a = owner_of_input(b)) a->unload(b); }
    long kernel_launches() const { long n = 0; for (const CudaAligner *a : dev_) n += a->kernel_launches(); return n; }
    void set_verbose(int v) { for (CudaAligner *a : dev_) a->verbose = v; }
private:
    CudaAligner *&owner_of_input(ReadBatch &b) { return reinterpret_cast<CudaAligner *&>(b.dev_owner); }
    std::vector<CudaAligner *> dev_;
};

} // namespace bsb
