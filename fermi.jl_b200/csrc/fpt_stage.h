// Host-side staging of caller-owned arrays: pageable host memory -> pinned bounce slots -> device.
//
// A Julia `ccall` (or numpy through ctypes) hands over ordinary pageable arrays.  cudaMemcpyAsync from pageable memory is
// staged by the driver on one thread (~11 GB/s measured on the B200 boxes, profiles/r01_ao_route_nbf144.json), five times below
// the PCIe rate.  Here the copy into pinned memory is done by a small pool of host threads, piece by piece, into a ring of
// pinned slots; every filled slot is sent with its own cudaMemcpyAsync, so the host copy of piece n+1 overlaps the DMA of
// piece n and -- with several GPUs in one process -- the DMAs of different GPUs run concurrently on their own PCIe links.
#pragma once
#include <cuda_runtime.h>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace fpt {

class CopyPool {
public:
    ~CopyPool() { stop(); }
    int size() const { return nthreads_; }
    // n = total number of threads taking part in a copy, the caller included (n - 1 workers are spawned)
    void start(int n)
    {
        stop();
        nthreads_ = n < 1 ? 1 : n;
        quit_ = false;
        for (int t = 1; t < nthreads_; t++) workers_.emplace_back([this] { run(); });
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
        nthreads_ = 1;
    }
    // memcpy split over the pool; returns when every byte is in place
    void copy(void* dst, const void* src, size_t bytes)
    {
        const size_t min_slice = (size_t)256 << 10;
        int parts = (int)((bytes + min_slice - 1) / min_slice);
        if (parts > nthreads_) parts = nthreads_;
        if (parts <= 1) { memcpy(dst, src, bytes); return; }
        size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int p = 1; p < parts; p++) {
                const size_t b = (size_t)p * slice;
                if (b >= bytes) break;
                const size_t n = bytes - b < slice ? bytes - b : slice;
                q_.push_back(Task{(char*)dst + b, (const char*)src + b, n});
                pending_++;
            }
        }
        cv_.notify_all();
        memcpy(dst, src, slice < bytes ? slice : bytes);
        std::unique_lock<std::mutex> lk(m_);
        // help with whatever is still queued, then wait for the stragglers
        while (!q_.empty()) {
            Task t = q_.front();
            q_.pop_front();
            lk.unlock();
            memcpy(t.dst, t.src, t.n);
            lk.lock();
            pending_--;
        }
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    struct Task { char* dst; const char* src; size_t n; };
    void run()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_.wait(lk, [this] { return quit_ || !q_.empty(); });
            if (quit_) return;
            Task t = q_.front();
            q_.pop_front();
            lk.unlock();
            memcpy(t.dst, t.src, t.n);
            lk.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<Task> q_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    int pending_ = 0;
    int nthreads_ = 1;
    bool quit_ = false;
};

// Ring of pinned bounce slots shared by the GPUs of one handle.  A slot is reusable once the DMA that read it has finished
// (an event recorded on the copy stream of the GPU it went to; events are tied to a device, so a slot keeps one per GPU).
struct PinnedRing {
    static constexpr size_t SLOT_BYTES = (size_t)4 << 20;
    struct Slot { char* p = nullptr; std::vector<cudaEvent_t> ev; int busy = -1; };
    std::vector<Slot> slots;
    size_t next = 0;
    cudaError_t init(int nslots, int ndev)
    {
        slots.resize(nslots);
        for (auto& s : slots) {
            s.ev.assign(ndev, nullptr);
            cudaError_t e = cudaHostAlloc((void**)&s.p, SLOT_BYTES, cudaHostAllocPortable);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    void release()
    {
        for (auto& s : slots) {
            for (cudaEvent_t e : s.ev)
                if (e) cudaEventDestroy(e);
            if (s.p) cudaFreeHost(s.p);
        }
        slots.clear();
    }
    // next slot, free to be overwritten by the host
    cudaError_t acquire(Slot** out)
    {
        Slot& s = slots[next];
        next = (next + 1) % slots.size();
        if (s.busy >= 0) {
            cudaError_t e = cudaEventSynchronize(s.ev[s.busy]);
            if (e != cudaSuccess) return e;
            s.busy = -1;
        }
        *out = &s;
        return cudaSuccess;
    }
    // the slot's content is being read by a DMA enqueued on `stream` of local GPU `idev` (the current device)
    cudaError_t sent(Slot* s, int idev, cudaStream_t stream)
    {
        if (!s->ev[idev]) {
            cudaError_t e = cudaEventCreateWithFlags(&s->ev[idev], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaEventRecord(s->ev[idev], stream);
        if (e == cudaSuccess) s->busy = idev;
        return e;
    }
};

}  // namespace fpt
