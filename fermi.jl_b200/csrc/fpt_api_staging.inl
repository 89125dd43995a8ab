// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: where a caller's pointer lives; host -> device staging through the pinned ring; `distribute` (sharded upload + all-gather).

// ---- where does a caller's pointer live ------------------------------------------------------------------------------------
enum PtrKind { PK_PAGEABLE = 0, PK_PINNED = 1, PK_DEVICE = 2 };
static PtrKind classify(const void* p, int* device = nullptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return PK_PAGEABLE; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        if (device) *device = at.device;
        return PK_DEVICE;
    }
    return at.type == cudaMemoryTypeHost ? PK_PINNED : PK_PAGEABLE;
}

// Device-resident inputs are consumed in place on the handle's own streams.  They must live on the handle's (first) GPU, and
// whatever stream produced them is ordered before the first read by one device-wide synchronisation.
static int admit_device_inputs(fpt_handle* h, const char* who, std::initializer_list<const void*> ptrs)
{
    bool any = false;
    for (const void* p : ptrs) {
        int dev = -1;
        if (p && classify(p, &dev) == PK_DEVICE) {
            if (dev != h->devs[0]->dev)
                return fail("%s: a device-resident input lives on GPU %d, the handle's GPU is %d", who, dev, h->devs[0]->dev);
            any = true;
        }
    }
    if (any) {
        CK(cudaSetDevice(h->devs[0]->dev));
        CK(cudaDeviceSynchronize());
    }
    return 0;
}

// ---- staging ---------------------------------------------------------------------------------------------------------------
// enqueue the copy of bytes [begin, begin + bytes) of the packed stream of `view` (a view of the host array `src`, see fpt_stage.h)
// to `dst` on GPU d (its copy stream).  Pinned or tiny contiguous sources are handed to the DMA engine directly, a pinned single
// slab of rows as a 2-D copy; everything else becomes a job for the staging threads, collected in `jobs` so that the parts of one
// array that go to different GPUs are staged as ONE transfer (stage_flush).  `view` must stay alive until the flush.
static int stage_to(fpt_handle* h, Dev& d, void* dst, const void* src, const View& view, size_t begin, size_t bytes, PtrKind kind,
                    std::vector<StagePool::Job>& jobs)
{
    if (bytes == 0) return 0;
    CK(cudaSetDevice(d.dev));
    h->h2d += (double)bytes;
    const bool one = view.slabs.size() == 1;
    const Slab& s0 = view.slabs[0];
    if (one && s0.nrows == 1 && (kind == PK_PINNED || bytes <= ((size_t)64 << 10))) {
        CK(cudaMemcpyAsync(dst, (const char*)src + s0.src_off + begin, bytes, cudaMemcpyHostToDevice, d.copy));
        return 0;
    }
    if (one && kind == PK_PINNED && begin % s0.row_bytes == 0 && bytes % s0.row_bytes == 0) {   // the DMA engine gathers the rows itself
        CK(cudaMemcpy2DAsync(dst, s0.row_bytes, (const char*)src + s0.src_off + (begin / s0.row_bytes) * s0.pitch, s0.pitch, s0.row_bytes,
                             bytes / s0.row_bytes, cudaMemcpyHostToDevice, d.copy));
        return 0;
    }
    StagePool::Job job;
    job.dst = (char*)dst; job.src = (const char*)src; job.view = &view; job.begin = begin; job.bytes = bytes;
    job.dev = d.dev; job.idev = d.idx; job.stream = d.copy;
    jobs.push_back(job);
    return 0;
}
static int stage_flush(fpt_handle* h, std::vector<StagePool::Job>& jobs)
{
    if (jobs.empty()) return 0;
    const auto t0 = wall::now();
    CK(h->pool.transfer(jobs));
    h->stage_host_ms += ms_since(t0);
    jobs.clear();
    return 0;
}
// one contiguous array to one GPU, staged right away
static int stage_now(fpt_handle* h, Dev& d, void* dst, const void* src, size_t bytes, PtrKind kind)
{
    std::vector<StagePool::Job> jobs;
    const View view = View::contiguous(bytes);
    if (stage_to(h, d, dst, src, view, 0, bytes, kind, jobs)) return 1;
    return stage_flush(h, jobs);
}

// the copy stream's work so far is what the compute stream continues from
static int copy_then_stream(Dev& d)
{
    CK(cudaSetDevice(d.dev));
    CK(cudaEventRecord(d.ev_copy, d.copy));
    CK(cudaStreamWaitEvent(d.stream, d.ev_copy, 0));
    return 0;
}

constexpr size_t SHARD_MIN_BYTES = (size_t)1 << 20;

// Makes the array `src` (host or device memory) resident on every GPU of the handle, ordered on each GPU's compute stream;
// out[i] = its address on devs[i].  `bufof(d)` names the staging buffer to use on GPU d.  `view` (host memory only): what of the array
// is wanted, in bytes (fpt_stage.h) -- its packed stream of n doubles is what arrives; nullptr = the n doubles at `src` themselves.
//  * host memory, world > 1, >= 1 MB: GPU g pulls only part g of `world` of the packed stream over its own PCIe link, then one
//    in-place ncclAllGather over NVLink completes it everywhere (in rank mode every process passes the same array and pulls its part);
//  * host memory otherwise: every GPU of this process pulls the whole stream;
//  * device memory (on devs[0]): used in place; the other GPUs of a single-process handle receive it by ncclBroadcast.
template <class BufOf>
static int distribute(fpt_handle* h, BufOf bufof, const double* src, size_t n, std::vector<const double*>& out, const View* view = nullptr)
{
    const int L = (int)h->devs.size(), W = h->world;
    out.assign(L, nullptr);
    const PtrKind kind = classify(src);
    if (kind == PK_DEVICE) {
        if (view) return fail("internal: views of device-resident arrays are not supported");
        out[0] = src;
        if (L > 1) {
            for (int i = 1; i < L; i++) {
                Dev& d = *h->devs[i];
                CK(cudaSetDevice(d.dev));
                if (bufof(d).ensure(n * sizeof(double))) return 1;
                out[i] = bufof(d).d();
            }
            NCK(nccl_api().GroupStart());
            for (int i = 0; i < L; i++) {
                Dev& d = *h->devs[i];
                NCK(nccl_api().Broadcast(out[i], (void*)out[i], n, ncclDouble, 0, d.comm, d.stream));
            }
            NCK(nccl_api().GroupEnd());
        }
        return 0;
    }
    const View whole = View::contiguous(n * sizeof(double));
    const View& vw = view ? *view : whole;
    if (vw.total != n * sizeof(double)) return fail("internal: view of %zu bytes for %zu doubles", vw.total, n);
    const bool shard = W > 1 && n * sizeof(double) >= SHARD_MIN_BYTES;
    std::vector<StagePool::Job> jobs;
    if (!shard) {
        for (int i = 0; i < L; i++) {
            Dev& d = *h->devs[i];
            CK(cudaSetDevice(d.dev));
            if (bufof(d).ensure(n * sizeof(double))) return 1;
            out[i] = bufof(d).d();
            if (stage_to(h, d, bufof(d).p, src, vw, 0, n * sizeof(double), kind, jobs)) return 1;
        }
        if (stage_flush(h, jobs)) return 1;
        for (int i = 0; i < L; i++)
            if (copy_then_stream(*h->devs[i])) return 1;
        return 0;
    }
    const size_t part = ((n + W - 1) / W + 511) & ~(size_t)511;
    for (int i = 0; i < L; i++) {
        Dev& d = *h->devs[i];
        CK(cudaSetDevice(d.dev));
        if (bufof(d).ensure((size_t)W * part * sizeof(double))) return 1;
        out[i] = bufof(d).d();
        const size_t b = std::min(n, (size_t)d.grank * part), e = std::min(n, (size_t)(d.grank + 1) * part);
        if (stage_to(h, d, bufof(d).d() + b, src, vw, b * sizeof(double), (e - b) * sizeof(double), kind, jobs)) return 1;
    }
    if (stage_flush(h, jobs)) return 1;
    for (int i = 0; i < L; i++)
        if (copy_then_stream(*h->devs[i])) return 1;
    NCK(nccl_api().GroupStart());
    for (int i = 0; i < L; i++) {
        Dev& d = *h->devs[i];
        NCK(nccl_api().AllGather(bufof(d).d() + (size_t)d.grank * part, bufof(d).p, part, ncclDouble, d.comm, d.stream));
    }
    NCK(nccl_api().GroupEnd());
    return 0;
}
