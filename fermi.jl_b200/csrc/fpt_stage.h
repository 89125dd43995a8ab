// Host-side staging of caller-owned arrays: pageable host memory -> pinned bounce slots -> device.
//
// A Julia `ccall` (or numpy through ctypes) hands over ordinary pageable arrays.  cudaMemcpyAsync from pageable memory is
// staged by the driver on one thread (~11 GB/s measured on the B200 boxes, profiles/r01_ao_route_nbf144.json), five times below
// the PCIe rate.  Here a pool of host threads does it: a transfer is cut into pieces of 2 MB, thread t takes pieces
// t, t + T, ...; for each it copies the piece into one of its own two pinned slots (non-temporal stores: the data is read next by
// the DMA engine, not by the CPU) and enqueues the slot's cudaMemcpyAsync itself.  No barrier per piece -- the threads only meet
// once per transfer -- and the copy of piece n+1 overlaps the DMA of piece n.  With several GPUs in one process the transfers to
// different GPUs are issued back to back, so their DMAs run concurrently on their own PCIe links.
#pragma once
#include <cuda_runtime.h>
#include <immintrin.h>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace fpt {

// memcpy with non-temporal stores (dst 32-byte aligned); falls back to memcpy on CPUs without AVX2
__attribute__((target("avx2"))) inline void copy_stream_avx2(void* dst, const void* src, size_t bytes)
{
    char* d = (char*)dst;
    const char* s = (const char*)src;
    size_t n = bytes / 128;
    for (size_t i = 0; i < n; i++) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s)), b = _mm256_loadu_si256((const __m256i*)(s + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(s + 64)), e = _mm256_loadu_si256((const __m256i*)(s + 96));
        _mm256_stream_si256((__m256i*)(d), a);
        _mm256_stream_si256((__m256i*)(d + 32), b);
        _mm256_stream_si256((__m256i*)(d + 64), c);
        _mm256_stream_si256((__m256i*)(d + 96), e);
        s += 128;
        d += 128;
    }
    if (bytes % 128) memcpy(d, s, bytes % 128);
    _mm_sfence();
}
inline void copy_to_pinned(void* dst, const void* src, size_t bytes, bool nt)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (nt && avx2 && ((uintptr_t)dst % 32) == 0) copy_stream_avx2(dst, src, bytes);
    else memcpy(dst, src, bytes);
}
// gather `nrows` rows of `row_bytes` (source pitch `pitch`) into a contiguous pinned buffer
__attribute__((target("avx2"))) inline void copy_rows_stream_avx2(char* d, const char* s, size_t nrows, size_t row_bytes, size_t pitch)
{
    for (size_t r = 0; r < nrows; r++, s += pitch) {
        for (size_t b = 0; b < row_bytes; b += 32, d += 32)
            _mm256_stream_si256((__m256i*)d, _mm256_loadu_si256((const __m256i*)(s + b)));
    }
    _mm_sfence();
}
inline void copy_rows_to_pinned(void* dst, const void* src, size_t nrows, size_t row_bytes, size_t pitch, bool nt)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (nt && avx2 && ((uintptr_t)dst % 32) == 0 && row_bytes % 32 == 0) { copy_rows_stream_avx2((char*)dst, (const char*)src, nrows, row_bytes, pitch); return; }
    char* d = (char*)dst;
    const char* s = (const char*)src;
    for (size_t r = 0; r < nrows; r++, s += pitch, d += row_bytes) memcpy(d, s, row_bytes);
}

class StagePool {
public:
    size_t PIECE = (size_t)2 << 20;   // bytes per piece = per pinned slot (set before start())
    static constexpr int SLOTS_PER_THREAD = 2;
    struct Job {
        char* dst = nullptr;          // device
        const char* src = nullptr;    // pageable host
        size_t bytes = 0;             // total bytes that arrive at dst (contiguous there)
        size_t row_bytes = 0;         // 0: `src` is contiguous too; else the source is a set of rows of row_bytes ...
        size_t src_pitch = 0;         // ... src_pitch bytes apart (a sub-range of the fastest index of a column-major array)
        int dev = 0;                  // CUDA ordinal
        int idev = 0;                 // index of the GPU inside the handle (event bank)
        cudaStream_t stream = nullptr;
    };
    ~StagePool() { stop(); }
    int size() const { return nthreads_; }
    bool nt_stores = true;

    // n = total number of threads taking part in a transfer, the caller included; ndev = GPUs of the handle
    cudaError_t start(int n, int ndev)
    {
        stop();
        nthreads_ = n < 1 ? 1 : n;
        ndev_ = ndev;
        slots_.assign((size_t)nthreads_ * SLOTS_PER_THREAD, Slot{});
        for (auto& s : slots_) {
            s.ev.assign(ndev, nullptr);
            cudaError_t e = cudaHostAlloc((void**)&s.p, PIECE, cudaHostAllocPortable);   // rows never exceed PIECE (o * 8 bytes)
            if (e != cudaSuccess) return e;
        }
        next_.assign(nthreads_, 0);
        quit_ = false;
        gen_ = 0;
        for (int t = 1; t < nthreads_; t++) workers_.emplace_back([this, t] { worker(t); });
        return cudaSuccess;
    }
    void stop()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
        for (auto& s : slots_) {
            for (cudaEvent_t e : s.ev)
                if (e) cudaEventDestroy(e);
            if (s.p) cudaFreeHost(s.p);
        }
        slots_.clear();
        nthreads_ = 1;
    }
    // enqueue the whole transfer (returns when every piece has been copied out of `src` and its DMA is enqueued)
    cudaError_t transfer(const Job& job)
    {
        const size_t np = (job.bytes + piece_bytes(job) - 1) / piece_bytes(job);
        const int active = (int)(np < (size_t)nthreads_ ? np : (size_t)nthreads_);
        err_.store((int)cudaSuccess);
        if (active > 1) {
            std::lock_guard<std::mutex> lk(m_);
            job_ = job;
            active_ = active;
            pending_ = active - 1;
            gen_++;
        } else {
            job_ = job;
            active_ = 1;
        }
        if (active > 1) cv_.notify_all();
        work(0);
        if (active > 1) {
            std::unique_lock<std::mutex> lk(m_);
            done_.wait(lk, [this] { return pending_ == 0; });
        }
        return (cudaError_t)err_.load();
    }

private:
    // bytes per piece: PIECE, or the whole number of rows that fits into it
    size_t piece_bytes(const Job& j) const { return j.row_bytes ? (PIECE / j.row_bytes ? PIECE / j.row_bytes : 1) * j.row_bytes : PIECE; }
    struct Slot { char* p = nullptr; std::vector<cudaEvent_t> ev; int busy = -1; };
    void note(cudaError_t e) { if (e != cudaSuccess) { int ok = (int)cudaSuccess; err_.compare_exchange_strong(ok, (int)e); } }
    void work(int t)
    {
        const Job job = job_;
        const int T = active_;
        if (t >= T) return;
        if (cudaSetDevice(job.dev) != cudaSuccess) { note(cudaGetLastError()); return; }
        const size_t pb = piece_bytes(job);
        const size_t np = (job.bytes + pb - 1) / pb;
        for (size_t p = t; p < np; p += T) {
            Slot& s = slots_[(size_t)t * SLOTS_PER_THREAD + next_[t]];
            next_[t] = (next_[t] + 1) % SLOTS_PER_THREAD;
            if (s.busy >= 0) {   // the DMA that last read this slot must have finished
                note(cudaEventSynchronize(s.ev[s.busy]));
                s.busy = -1;
            }
            const size_t off = p * pb, nb = job.bytes - off < pb ? job.bytes - off : pb;
            if (job.row_bytes) copy_rows_to_pinned(s.p, job.src + (off / job.row_bytes) * job.src_pitch, nb / job.row_bytes, job.row_bytes, job.src_pitch, nt_stores);
            else copy_to_pinned(s.p, job.src + off, nb, nt_stores);
            note(cudaMemcpyAsync(job.dst + off, s.p, nb, cudaMemcpyHostToDevice, job.stream));
            if (!s.ev[job.idev]) note(cudaEventCreateWithFlags(&s.ev[job.idev], cudaEventDisableTiming));
            if (s.ev[job.idev]) {
                note(cudaEventRecord(s.ev[job.idev], job.stream));
                s.busy = job.idev;
            }
        }
    }
    void worker(int t)
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
            if (quit_) return;
            seen = gen_;
            if (t >= active_) continue;
            lk.unlock();
            work(t);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::vector<Slot> slots_;
    std::vector<int> next_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::atomic<int> err_{0};
    Job job_;
    unsigned long long gen_ = 0;
    int active_ = 1, pending_ = 0;
    int nthreads_ = 1, ndev_ = 1;
    bool quit_ = false;
};

}  // namespace fpt
