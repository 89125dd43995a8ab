// Host-side staging of caller-owned arrays: pageable host memory -> pinned bounce slots -> device.
//
// A Julia `ccall` (or numpy through ctypes) hands over ordinary pageable arrays.  cudaMemcpyAsync from pageable memory is
// staged by the driver on one thread (~11 GB/s measured on the B200 boxes, profiles/r01_ao_route_nbf144.json), five times below
// the PCIe rate.  Here a pool of host threads does it: a transfer is cut into pieces of 2 MB, thread t takes pieces
// t, t + T, ...; for each it copies the piece into one of its own two pinned slots (non-temporal stores: the data is read next by
// the DMA engine, not by the CPU) and enqueues the slot's cudaMemcpyAsync itself.  No barrier per piece -- the threads only meet
// once per transfer -- and the copy of piece n+1 overlaps the DMA of piece n.  With several GPUs in one process the parts of an
// array that go to different GPUs form ONE transfer whose pieces are dealt round-robin over the GPUs, so all PCIe links are
// busy at once and the threads still meet only once per array.
//
// What is copied is a *view* of the caller's array: a list of slabs, each a run of equally long rows a fixed pitch apart.  A plain
// array is one slab of one row; a slice of an array's fastest index is one slab of short rows; the symmetry-unique half of
// OVVV[i,a,b,c] = OVVV[i,a,c,b] is one slab per c holding the prefix b <= c.  The view's rows arrive back to back ("packed") on the
// device, and a transfer may cover any byte range of the packed stream (one GPU's share of a sharded upload).
#pragma once
#include <cuda_runtime.h>
#include <immintrin.h>
#include <algorithm>
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
// A strided view of host memory (all sizes in bytes)
struct Slab {
    size_t src_off;      // first row, relative to the array
    size_t nrows, row_bytes, pitch;
    size_t packed_off;   // where the slab starts in the packed stream
};
struct View {
    std::vector<Slab> slabs;
    size_t total = 0;
    void add(size_t src_off, size_t nrows, size_t row_bytes, size_t pitch)
    {
        if (nrows == 0 || row_bytes == 0) return;
        if (pitch == row_bytes) { row_bytes *= nrows; pitch = row_bytes; nrows = 1; }   // contiguous rows are one row
        slabs.push_back(Slab{src_off, nrows, row_bytes, pitch, total});
        total += nrows * row_bytes;
    }
    static View contiguous(size_t bytes) { View v; v.add(0, 1, bytes, bytes); return v; }
};

// The views the (T) uploads use (sizes in doubles; the packed layouts are what prep_pt_particle(_tri), expand_t2_tri and expand_ovov_tri
// in fpt_aux_kernels.cuh read -- tests/test_stage_views.py checks both sides of that contract on the CPU):
//   OVVV[i,a,b,c], c in [c0, c0+cn), occupied slice [p0, p0+np):  per c the rows (a, b) of np doubles, b < v -- or b <= c (half)
inline View view_ovvv_chunk(size_t o, size_t v, size_t p0, size_t np, size_t c0, size_t cn, bool half)
{
    View vw;
    for (size_t c = c0; c < c0 + cn; c++)
        vw.add(((c * v * v * o) + p0) * sizeof(double), v * (half ? c + 1 : v), np * sizeof(double), o * sizeof(double));
    return vw;
}
//   T2[i,j,a,b] = T2[j,i,b,a]: per b the prefix a <= b of the (i, j, a) block
inline View view_t2_half(size_t o, size_t v)
{
    View vw;
    for (size_t b = 0; b < v; b++) vw.add(b * o * o * v * sizeof(double), 1, o * o * (b + 1) * sizeof(double), o * o * (b + 1) * sizeof(double));
    return vw;
}
//   OVOV[i,a,j,b] = OVOV[j,b,i,a]: per (b, j) the prefix a <= b of the (i, a) plane
inline View view_ovov_half(size_t o, size_t v)
{
    View vw;
    for (size_t b = 0; b < v; b++) vw.add(b * o * v * o * sizeof(double), o, o * (b + 1) * sizeof(double), o * v * sizeof(double));
    return vw;
}

// copy bytes [off, off + nb) of the packed stream of `view` (over the array at `src`) to `dst` (a pinned slot, 32-byte aligned)
__attribute__((target("avx2"))) inline void copy_seg_avx2(char* d, const char* s, size_t n)
{
    size_t b = 0;
    for (; b + 32 <= n; b += 32) _mm256_stream_si256((__m256i*)(d + b), _mm256_loadu_si256((const __m256i*)(s + b)));
    if (b < n) memcpy(d + b, s + b, n - b);
}
inline void copy_view_to_pinned(char* dst, const char* src, const View& view, size_t off, size_t nb, bool nt)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    // slab that holds packed offset `off`
    size_t lo = 0, hi = view.slabs.size();
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) / 2;
        if (view.slabs[mid].packed_off <= off) lo = mid; else hi = mid;
    }
    char* d = dst;
    bool streamed = false;
    for (size_t si = lo; nb > 0 && si < view.slabs.size(); si++) {
        const Slab& sl = view.slabs[si];
        const size_t within = off - sl.packed_off;
        size_t row = within / sl.row_bytes, col = within % sl.row_bytes;
        for (; row < sl.nrows && nb > 0; row++, col = 0) {
            const size_t n = std::min(sl.row_bytes - col, nb);
            const char* from = src + sl.src_off + row * sl.pitch + col;
            if (nt && avx2 && ((uintptr_t)d % 32) == 0) {
                if (n >= 4096) copy_stream_avx2(d, from, n); else { copy_seg_avx2(d, from, n); streamed = true; }
            }
            else memcpy(d, from, n);
            d += n; off += n; nb -= n;
        }
    }
    if (streamed) _mm_sfence();
}

class StagePool {
public:
    size_t PIECE = (size_t)2 << 20;   // bytes per piece = per pinned slot (set before start())
    static constexpr int SLOTS_PER_THREAD = 2;
    struct Job {
        char* dst = nullptr;          // device address of the job's first byte
        const char* src = nullptr;    // the caller's array (pageable host memory)
        const View* view = nullptr;   // what of it is copied (must outlive the transfer)
        size_t begin = 0;             // packed range [begin, begin + bytes) of the view
        size_t bytes = 0;
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
            cudaError_t e = cudaHostAlloc((void**)&s.p, PIECE, cudaHostAllocPortable);
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
    // enqueue a set of transfers (returns when every piece has been copied out of its `src` and its DMA is enqueued)
    cudaError_t transfer(const std::vector<Job>& jobs)
    {
        // piece list: local piece lp of job j, jobs interleaved (lp slowest) so that concurrent pieces go to different GPUs
        plan_.clear();
        size_t maxnp = 0;
        for (const Job& j : jobs) maxnp = std::max(maxnp, npieces(j));
        for (size_t lp = 0; lp < maxnp; lp++)
            for (size_t j = 0; j < jobs.size(); j++)
                if (lp < npieces(jobs[j])) plan_.push_back({(int)j, lp});
        if (plan_.empty()) return cudaSuccess;
        const int active = (int)(plan_.size() < (size_t)nthreads_ ? plan_.size() : (size_t)nthreads_);
        err_.store((int)cudaSuccess);
        {
            std::lock_guard<std::mutex> lk(m_);
            jobs_ = &jobs;
            active_ = active;
            pending_ = active - 1;
            if (active > 1) gen_++;
        }
        if (active > 1) cv_.notify_all();
        work(0);
        if (active > 1) {
            std::unique_lock<std::mutex> lk(m_);
            done_.wait(lk, [this] { return pending_ == 0; });
        }
        return (cudaError_t)err_.load();
    }
    cudaError_t transfer(const Job& job) { const std::vector<Job> one(1, job); return transfer(one); }

private:
    struct Slot { char* p = nullptr; std::vector<cudaEvent_t> ev; int busy = -1; };
    void note(cudaError_t e) { if (e != cudaSuccess) { int ok = (int)cudaSuccess; err_.compare_exchange_strong(ok, (int)e); } }
    size_t npieces(const Job& j) const { return (j.bytes + PIECE - 1) / PIECE; }
    void work(int t)
    {
        const std::vector<Job>& jobs = *jobs_;
        const int T = active_;
        if (t >= T) return;
        int cur_dev = -1;
        for (size_t g = t; g < plan_.size(); g += T) {
            const Job& job = jobs[plan_[g].job];
            if (job.dev != cur_dev) {
                if (cudaSetDevice(job.dev) != cudaSuccess) { note(cudaGetLastError()); return; }
                cur_dev = job.dev;
            }
            const size_t pb = PIECE, p = plan_[g].piece;
            Slot& s = slots_[(size_t)t * SLOTS_PER_THREAD + next_[t]];
            next_[t] = (next_[t] + 1) % SLOTS_PER_THREAD;
            if (s.busy >= 0) {   // the DMA that last read this slot must have finished
                note(cudaEventSynchronize(s.ev[s.busy]));
                s.busy = -1;
            }
            const size_t off = p * pb, nb = job.bytes - off < pb ? job.bytes - off : pb;
            copy_view_to_pinned(s.p, job.src, *job.view, job.begin + off, nb, nt_stores);
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
    struct Piece { int job; size_t piece; };
    std::vector<Piece> plan_;
    const std::vector<Job>* jobs_ = nullptr;
    unsigned long long gen_ = 0;
    int active_ = 1, pending_ = 0;
    int nthreads_ = 1, ndev_ = 1;
    bool quit_ = false;
};

}  // namespace fpt
