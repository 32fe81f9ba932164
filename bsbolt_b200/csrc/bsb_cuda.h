// bsb_cuda.h -- the GPU batch aligner (implemented in bsb_cuda.cu; CUDA only, no CPU fallback)
#pragma once
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
    void preload(ReadBatch &b) override;
    void unload(ReadBatch &b) override;
    long kernel_launches() const;   // kernels launched by this object so far
    int device() const;
    size_t index_bytes() const;     // HBM held by the resident index
    int verbose = 3;
private:
    struct Impl;
    Impl *im_;
};

} // namespace bsb
