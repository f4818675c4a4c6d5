// hostpack.h -- host-side marshalling for the end-to-end path (hostpack.cpp): 2-bit packing of
// all-ACGT k-mers by a small worker pool.  Plain C++ (compiled by g++, no CUDA types).
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

namespace msbwt {

// Packs queries [q0, q1) of a batch of n_total fixed-length k-mers (`syms`, one symbol per byte) into
// out[w * stride + (q - qbase)], w < ceil(k/32): 2 bits per symbol, the k-mer's last symbol in the top
// bits of word 0.  Queries holding any symbol outside ACGT are appended to `exc` (their words are
// still written, with those symbols as 'A'; the caller must not use them).
void host_pack_range(const uint8_t *syms, uint32_t k, uint64_t n_total, uint64_t q0, uint64_t q1, uint64_t qbase,
                     uint64_t stride, uint64_t *out, std::vector<uint64_t> &exc);

// usable host threads for this process: the affinity mask, divided by LOCAL_WORLD_SIZE when several
// ranks share the host (torchrun), overridden by MSBWT_HOST_THREADS; at most 64
int host_threads_available();

// fork-join pool: between begin_session() and end_session() the workers spin, and run(fn) calls
// fn(tid, nthreads) on every worker (the caller is worker 0) and returns when all are done.  Sessions nest
// (the workers park when the outermost one ends); one run at a time; outside a session the workers sleep.
class HostPool {
  public:
    explicit HostPool(int nthreads);
    ~HostPool();
    HostPool(const HostPool &) = delete;
    HostPool &operator=(const HostPool &) = delete;
    int size() const;
    void begin_session();
    void end_session();
    void run(const std::function<void(int, int)> &fn);

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace msbwt
