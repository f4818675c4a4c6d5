// hostpack.cpp -- host side of the end-to-end path: a small worker pool that packs the caller's
// k-mers (one symbol per byte, the `&[u8]` the reference's count_kmer takes, src/msbwt_core.rs:125)
// into 2-bit words while they are staged for the copy to the device.
//
// Why: through the C ABI the batch arrives in host memory, and at one byte per symbol the PCIe link
// (~55 GB/s) caps a B200 at ~1.7 G 31-mers/s -- half of what the search kernel sustains.  The bytes
// have to be read once by the CPU anyway to reach the pinned staging buffer; packing them on the way
// cuts the copy to 8 bytes per 31-mer.  This is marshalling only: no rank, no search, no count is
// computed here, and a k-mer with any symbol outside ACGT (or >= 6) is not packed at all -- it is
// reported as an exception and travels as bytes to the device, which validates and counts it.
//
// Word format (matches seed_packed_kernel, kernels.cu): ceil(k/32) u64 words per k-mer, word-major
// (`out[w * stride + q]`); word w holds the symbols consumed at steps 32w .. 32w+31 of the backward
// search (the k-mer's LAST symbol is step 0), 2 bits each (A,C,G,T = 0..3), step 32w in the top bits.
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#include "hostpack.h"

namespace msbwt {

// ---------------------------------------------------------------- packing

namespace {

// code | 0x80 for a symbol outside ACGT (low nibble index)
alignas(32) const uint8_t kCodeLut[32] = {0x80, 0, 1, 2, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80,
                                          0x80, 0, 1, 2, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80};

// c (1..32) symbol bytes at p -> little-endian 2-bit pack (symbol i at bits 2i); *bad |= any symbol outside ACGT
inline uint64_t pack_chunk_scalar(const uint8_t *p, uint32_t c, bool *bad) {
    uint64_t w = 0;
    for (uint32_t i = 0; i < c; i++) {
        const uint8_t sy = p[i];
        const uint8_t code = sy < 16 ? kCodeLut[sy] : 0x80;
        if (code & 0x80) *bad = true;
        w |= (uint64_t)(code & 3u) << (2 * i);
    }
    return w;
}

__attribute__((target("avx2"))) inline uint64_t pack_chunk_avx2(const uint8_t *p, __m256i keep, bool *bad) {
    // bytes outside the chunk are replaced by 'A' (code 0, not an exception)
    const __m256i raw = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const __m256i x = _mm256_blendv_epi8(_mm256_set1_epi8(1), raw, keep);
    const __m256i lut = _mm256_load_si256(reinterpret_cast<const __m256i *>(kCodeLut));
    const __m256i code = _mm256_shuffle_epi8(lut, x);  // index bit 7 set -> 0: caught by `x` itself below
    const __m256i flags = _mm256_or_si256(_mm256_or_si256(code, x), _mm256_cmpgt_epi8(x, _mm256_set1_epi8(15)));
    if (_mm256_movemask_epi8(flags)) *bad = true;
    const __m256i c2 = _mm256_and_si256(code, _mm256_set1_epi8(3));
    const __m256i n4 = _mm256_maddubs_epi16(c2, _mm256_set1_epi16(0x0401));      // 2 symbols -> 4 bits
    const __m256i n8 = _mm256_madd_epi16(n4, _mm256_set1_epi32(0x00100001));     // 4 symbols -> 8 bits per dword
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i b = _mm256_shuffle_epi8(n8, pick);
    return (uint64_t)(uint32_t)_mm256_extract_epi32(b, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(b, 4) << 32);
}

__attribute__((target("avx2"))) __m256i keep_mask(uint32_t c) {  // 0xFF for byte index < c
    alignas(32) uint8_t m[32];
    for (uint32_t i = 0; i < 32; i++) m[i] = i < c ? 0xFF : 0;
    return _mm256_load_si256(reinterpret_cast<const __m256i *>(m));
}

__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t *syms, uint32_t k, uint64_t q0, uint64_t q1,
                                                     uint64_t qbase, uint64_t stride, uint64_t *out,
                                                     const uint8_t *safe_end, std::vector<uint64_t> &exc) {
    const uint32_t nw = (k + 31) / 32, tail = k - 32 * (nw - 1);  // symbols in the last word (1..32)
    const __m256i keep_full = keep_mask(32), keep_tail = keep_mask(tail);
    for (uint64_t q = q0; q < q1; q++) {
        const uint8_t *src = syms + q * k;
        bool bad = false;
        for (uint32_t w = 0; w < nw; w++) {
            const bool last = w + 1 == nw;
            const uint32_t c = last ? tail : 32;
            const uint8_t *p = last ? src : src + (k - 32 * (w + 1));
            uint64_t le;
            if (p <= safe_end) {
                le = pack_chunk_avx2(p, last ? keep_tail : keep_full, &bad);
            } else {  // a 32-byte load would run past the end of the caller's buffer
                le = pack_chunk_scalar(p, c, &bad);
            }
            out[(uint64_t)w * stride + (q - qbase)] = le << (64 - 2 * c);
        }
        if (bad) exc.push_back(q);
    }
}

void pack_range_scalar(const uint8_t *syms, uint32_t k, uint64_t q0, uint64_t q1, uint64_t qbase, uint64_t stride,
                       uint64_t *out, std::vector<uint64_t> &exc) {
    const uint32_t nw = (k + 31) / 32, tail = k - 32 * (nw - 1);
    for (uint64_t q = q0; q < q1; q++) {
        const uint8_t *src = syms + q * k;
        bool bad = false;
        for (uint32_t w = 0; w < nw; w++) {
            const bool last = w + 1 == nw;
            const uint32_t c = last ? tail : 32;
            const uint8_t *p = last ? src : src + (k - 32 * (w + 1));
            out[(uint64_t)w * stride + (q - qbase)] = pack_chunk_scalar(p, c, &bad) << (64 - 2 * c);
        }
        if (bad) exc.push_back(q);
    }
}

}  // namespace

void host_pack_range(const uint8_t *syms, uint32_t k, uint64_t n_total, uint64_t q0, uint64_t q1, uint64_t qbase,
                     uint64_t stride, uint64_t *out, std::vector<uint64_t> &exc) {
    if (!k || q0 >= q1) return;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const uint64_t bytes = n_total * (uint64_t)k;
    if (avx2 && bytes >= 32) pack_range_avx2(syms, k, q0, q1, qbase, stride, out, syms + (bytes - 32), exc);
    else pack_range_scalar(syms, k, q0, q1, qbase, stride, out, exc);
}

// ---------------------------------------------------------------- worker pool

// Workers sleep on a condition variable between sessions; inside a session (one batch call) they spin briefly
// on the job generation -- handing them a chunk then costs about a microsecond, not a wake-up -- and fall back to
// a timed sleep when no job shows up for a while (the issuing thread is waiting for the copy engine, or the host
// has fewer free cores than the pool has threads: spinning workers would then starve the packers that have work).
struct HostPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go;
    std::mutex job_mu;
    std::condition_variable cv_job;
    std::atomic<const std::function<void(int, int)> *> job{nullptr};
    std::atomic<uint64_t> generation{0};
    std::atomic<int> pending{0};
    std::atomic<int> sleepers{0};
    std::atomic<bool> in_session{false};
    int session_depth = 0;  // begin_session / end_session nest (guarded by mu)
    bool stop = false;
    int nthreads = 1;

    void worker(int tid) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || in_session.load(std::memory_order_acquire); });
                if (stop) return;
            }
            uint32_t idle = 0;
            while (in_session.load(std::memory_order_acquire)) {
                const uint64_t g = generation.load(std::memory_order_acquire);
                if (g == seen) {
                    if (++idle < 20000u) {
                        _mm_pause();
                        continue;
                    }
                    // nothing for ~0.1 ms: sleep until the next job (or at most 1 ms, to notice the session's end)
                    std::unique_lock<std::mutex> lk(job_mu);
                    sleepers.fetch_add(1, std::memory_order_acq_rel);
                    cv_job.wait_for(lk, std::chrono::milliseconds(1), [&] {
                        return generation.load(std::memory_order_acquire) != seen || !in_session.load(std::memory_order_acquire);
                    });
                    sleepers.fetch_sub(1, std::memory_order_acq_rel);
                    continue;
                }
                idle = 0;
                seen = g;
                (*job.load(std::memory_order_acquire))(tid, nthreads);
                pending.fetch_sub(1, std::memory_order_acq_rel);
            }
        }
    }
};

int host_threads_available() {
    int n = 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 1;
    // one process per GPU on a shared host (torchrun): split the cores between the local ranks
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) {
        const int w = atoi(lw);
        if (w > 1) n = n / w > 0 ? n / w : 1;
    }
    if (const char *env = getenv("MSBWT_HOST_THREADS")) {
        const int v = atoi(env);
        if (v > 0) n = v;
    }
    return n > 64 ? 64 : n;
}

HostPool::HostPool(int nthreads) : impl_(new Impl) {
    impl_->nthreads = nthreads < 1 ? 1 : nthreads;
    for (int t = 1; t < impl_->nthreads; t++) impl_->workers.emplace_back([this, t] { impl_->worker(t); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv_go.notify_all();
    for (auto &w : impl_->workers) w.join();
    delete impl_;
}

int HostPool::size() const { return impl_->nthreads; }

void HostPool::begin_session() {
    if (impl_->nthreads == 1) return;
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        if (impl_->session_depth++ > 0) return;  // nested: the workers are awake already
        impl_->in_session.store(true, std::memory_order_release);
    }
    impl_->cv_go.notify_all();
}

void HostPool::end_session() {
    if (impl_->nthreads == 1) return;
    std::lock_guard<std::mutex> lk(impl_->mu);
    if (--impl_->session_depth > 0) return;
    impl_->session_depth = 0;
    impl_->in_session.store(false, std::memory_order_release);
}

// inside a session only
void HostPool::run(const std::function<void(int, int)> &fn) {
    if (impl_->nthreads == 1) { fn(0, 1); return; }
    impl_->job.store(&fn, std::memory_order_release);
    impl_->pending.store(impl_->nthreads - 1, std::memory_order_release);
    impl_->generation.fetch_add(1, std::memory_order_acq_rel);
    if (impl_->sleepers.load(std::memory_order_acquire) > 0) {
        std::lock_guard<std::mutex> lk(impl_->job_mu);
        impl_->cv_job.notify_all();
    }
    fn(0, impl_->nthreads);  // the caller is worker 0
    while (impl_->pending.load(std::memory_order_acquire) != 0) _mm_pause();
}

}  // namespace msbwt
