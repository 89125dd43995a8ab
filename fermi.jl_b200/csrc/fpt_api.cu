// C-ABI of libfermi_pt_b200.so (declared in include/fermi_pt_b200.h) and the host-side driver of the (T) path:
// device buffer pool, sharded host->device staging (pinned bounce ring, one PCIe link per GPU), NCCL all-gather of the raw
// arrays over NVLink, layout prep / DF assembly launches, the persistent fused kernel, the scalar all-reduce of E(T).
// Replaces the driver loop + accumulation of
//   src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150 (reference, Julia threads)
// No CPU fallback: every entry point fails loudly without a CUDA device.
#include <cuda_runtime.h>
#include <sched.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fpt_internal.h"
#include "fpt_aux_kernels.cuh"
#include "fpt_gemm.cuh"
#include "fpt_triples.cuh"
#ifdef FPT_WITH_VARIANT2
#include "fpt_triples2.cuh"
#endif

using namespace fpt;
typedef std::chrono::steady_clock wall;
static double ms_since(wall::time_point t0) { return std::chrono::duration<double, std::milli>(wall::now() - t0).count(); }

extern "C" const char* fpt_last_error(void) { return err_text().c_str(); }
extern "C" const char* fpt_version(void) { return "fermi_pt_b200 0.4 (sm_100a)"; }

// ---- handle life cycle -----------------------------------------------------------------------------------------------------
// every device buffer of one GPU of a handle
static std::vector<DevBuf*> all_bufs(Dev* d)
{
    std::vector<DevBuf*> v = {&d->Pt, &d->Qt, &d->OV2, &d->T1d, &d->fo, &d->fv, &d->partials, &d->counter, &d->out, &d->prof, &d->blocktab,
                              &d->sT1, &d->sT2, &d->sOOOV, &d->sOVOV, &d->sChunk[0], &d->sChunk[1], &d->sTri, &d->sTri2, &d->sBOO, &d->sBOV, &d->sBVV,
                              &d->xTau, &d->xSlab, &d->xNew, &d->xOVOV, &d->ringtab,
                              &d->sCo, &d->sCv, &d->aoDense, &d->sIdx, &d->sVals, &d->aoQ1, &d->aoQ2v, &d->aoQ2o, &d->aoQ3vv, &d->aoQ3vo,
                              &d->aoQ3oo, &d->aoOVVV, &d->aoOOOV, &d->aoOVOV, &d->aoFlag};
    for (DevBuf& b : d->sPhase) v.push_back(&b);
    for (DevBuf& b : d->f32in) v.push_back(&b);
    for (DevBuf& b : d->f32wide) v.push_back(&b);
    return v;
}

static void dev_destroy(Dev* d)
{
    if (!d) return;
    cudaSetDevice(d->dev);
    if (d->comm) nccl_api().CommDestroy(d->comm);
    const std::vector<DevBuf*> bufs = all_bufs(d);
    for (DevBuf* b : bufs) b->release();
    cudaEvent_t evs[] = {d->ev0[0], d->ev1[0], d->ev0[1], d->ev1[1], d->ev0[2], d->ev1[2], d->ev0[3], d->ev1[3], d->ev_copy, d->ev_start, d->ev_free[0], d->ev_free[1]};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d->tl)
        if (e) cudaEventDestroy(e);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->copy) cudaStreamDestroy(d->copy);
    delete d;
}

static int dev_init(Dev* d)
{
    cudaDeviceProp prop;
    CK(cudaSetDevice(d->dev));
    CK(cudaGetDeviceProperties(&prop, d->dev));
    if (prop.major < 10)
        return fail("fpt_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", d->dev, prop.major, prop.minor);
    d->n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking));
    static_assert(MAX_PHASES == 4, "dev_destroy lists the phase events and buffers one by one");
    for (int t = 0; t < MAX_PHASES; t++) {
        CK(cudaEventCreate(&d->ev0[t]));
        CK(cudaEventCreate(&d->ev1[t]));
    }
    CK(cudaEventCreateWithFlags(&d->ev_copy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_free[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_free[1], cudaEventDisableTiming));
    for (int t = 0; t < NTL; t++) CK(cudaEventCreate(&d->tl[t]));
    CK(cudaFuncSetAttribute(triples_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
#ifdef FPT_WITH_VARIANT2
    CK(cudaFuncSetAttribute(triples_kernel2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
#endif
    CK(gemm_tn_set_attributes<EPI_COLMAJOR>());
    CK(gemm_tn_set_attributes<EPI_PT>());
    CK(gemm_tn_set_attributes<EPI_QT_HOLE>());
    CK(gemm_tn_set_attributes<EPI_OV2>());
    CK(gemm_tn_set_attributes<EPI_LADDER_SLAB>());
    CK(gemm_tn_set_attributes<EPI_LADDER_OUT>());
    return 0;
}

static int dev_create(int dev, int idx, Dev** out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("fpt_create: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail("fpt_create: device %d out of range (have %d)", dev, ndev);
    Dev* d = new Dev();
    d->dev = dev;
    d->idx = idx;
    if (dev_init(d)) { dev_destroy(d); return 1; }
    *out = d;
    return 0;
}

static int default_host_threads(int share)
{
    if (const char* s = getenv("FERMI_PT_B200_THREADS")) {
        const int n = atoi(s);
        if (n >= 1) return n > 64 ? 64 : n;
    }
    int n = 1;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    n /= (share > 0 ? share : 1);
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

// pool of staging threads, pinned bounce ring, pinned result word
static int handle_finish(fpt_handle* h)
{
    CK(cudaSetDevice(h->devs[0]->dev));
    if (const char* nt = getenv("FERMI_PT_B200_NT")) h->pool.nt_stores = atoi(nt) != 0;
    if (const char* kb = getenv("FERMI_PT_B200_PIECE_KB")) {
        const long n = atol(kb);
        if (n >= 64 && n <= 65536) h->pool.PIECE = (size_t)n << 10;
    }
    CK(h->pool.start(default_host_threads(h->rank_mode ? h->world : 1), (int)h->devs.size()));
    CK(cudaHostAlloc((void**)&h->res_pinned, 64, cudaHostAllocPortable));
    return 0;
}

extern "C" int fpt_destroy(fpt_handle* h)
{
    if (!h) return 0;
    DeviceGuard guard;
    for (Dev* d : h->devs) {
        cudaSetDevice(d->dev);
        cudaDeviceSynchronize();
    }
    h->pool.stop();
    if (h->res_pinned) cudaFreeHost(h->res_pinned);
    for (Dev* d : h->devs) dev_destroy(d);
    delete h;
    return 0;
}

extern "C" int fpt_create(int ngpu, const int* devices, fpt_handle** out)
{
    if (!out) return fail("fpt_create: out is NULL");
    *out = nullptr;
    if (ngpu < 1 || ngpu > 16) return fail("fpt_create: ngpu=%d out of range", ngpu);
    if (ngpu > 1 && !devices) return fail("fpt_create: a device list is required for ngpu > 1");
    DeviceGuard guard;
    fpt_handle* h = new fpt_handle();
    for (int k = 0; k < ngpu; k++) {
        Dev* d = nullptr;
        if (dev_create(devices ? devices[k] : -1, k, &d)) { fpt_destroy(h); return 1; }
        d->grank = k;
        h->devs.push_back(d);
    }
    h->world = ngpu;
    if (ngpu > 1) {
        // single-process multi-GPU: one NCCL clique over NVLink
        if (nccl_load()) { fpt_destroy(h); return 1; }
        std::vector<ncclComm_t> comms(ngpu);
        std::vector<int> devs(ngpu);
        for (int k = 0; k < ngpu; k++) devs[k] = h->devs[k]->dev;
        ncclResult_t r = nccl_api().CommInitAll(comms.data(), ngpu, devs.data());
        if (r != ncclSuccess) {
            fail("ncclCommInitAll failed: %s", nccl_api().GetErrorString(r));
            fpt_destroy(h);
            return 1;
        }
        for (int k = 0; k < ngpu; k++) h->devs[k]->comm = comms[k];
    }
    if (handle_finish(h)) { fpt_destroy(h); return 1; }
    *out = h;
    return 0;
}

extern "C" int fpt_nccl_unique_id(void* id128)
{
    if (!id128) return fail("fpt_nccl_unique_id: NULL argument");
    if (nccl_load()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI carries the NCCL id as 128 bytes");
    ncclUniqueId id;
    NCK(nccl_api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return 0;
}

extern "C" int fpt_create_rank(int device, int rank, int world, const void* id128, fpt_handle** out)
{
    if (!out) return fail("fpt_create_rank: out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail("fpt_create_rank: invalid rank %d of %d", rank, world);
    if (world > 1 && !id128) return fail("fpt_create_rank: the NCCL id is required for world > 1");
    DeviceGuard guard;
    fpt_handle* h = new fpt_handle();
    Dev* d = nullptr;
    if (dev_create(device, 0, &d)) { fpt_destroy(h); return 1; }
    d->grank = rank;
    h->devs.push_back(d);
    h->world = world;
    h->rank_mode = true;
    if (world > 1) {
        if (nccl_load()) { fpt_destroy(h); return 1; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        cudaSetDevice(d->dev);
        ncclResult_t r = nccl_api().CommInitRank(&d->comm, world, id, rank);
        if (r != ncclSuccess) {
            d->comm = nullptr;
            fail("ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
            fpt_destroy(h);
            return 1;
        }
    }
    if (handle_finish(h)) { fpt_destroy(h); return 1; }
    *out = h;
    return 0;
}

extern "C" int fpt_set_symmetric_inputs(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_symmetric_inputs: NULL handle");
    h->sym_inputs = on ? 1 : 0;
    return 0;
}

extern "C" int fpt_set_deterministic(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_deterministic: NULL handle");
    h->deterministic = on ? 1 : 0;
    return 0;
}

extern "C" int fpt_set_df_ring(fpt_handle* h, int block)
{
    if (!h) return fail("fpt_set_df_ring: NULL handle");
    if (block < -1 || block > 64) return fail("fpt_set_df_ring: block=%d out of range (-1 .. 64)", block);
    h->df_ring = block;
    return 0;
}

extern "C" int fpt_device_bytes(fpt_handle* h, double* bytes)
{
    if (!h || !bytes) return fail("fpt_device_bytes: NULL argument");
    const std::vector<DevBuf*> bufs = all_bufs(h->devs[0]);
    double n = 0.0;
    for (DevBuf* b : bufs) n += (double)b->cap;
    *bytes = n;
    return 0;
}

extern "C" int fpt_set_host_threads(fpt_handle* h, int n)
{
    if (!h) return fail("fpt_set_host_threads: NULL handle");
    if (n < 1 || n > 64) return fail("fpt_set_host_threads: n=%d out of range (1..64)", n);
    DeviceGuard guard;
    for (Dev* d : h->devs) {   // no DMA may still be reading the slots that are about to be freed
        CK(cudaSetDevice(d->dev));
        CK(cudaStreamSynchronize(d->copy));
    }
    CK(h->pool.start(n, (int)h->devs.size()));
    return 0;
}

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

// ---- problem set-up --------------------------------------------------------------------------------------------------------
static int grid1d(i64 n, int block = 256) { i64 g = (n + block - 1) / block; if (g > 148 * 32) g = 148 * 32; if (g < 1) g = 1; return (int)g; }

static int check_idle(fpt_handle* h, const char* who)
{
    if (!h) return fail("%s: NULL handle", who);
    if (h->pending) return fail("%s: an asynchronous call is in flight on this handle; collect it with fpt_wait first", who);
    return 0;
}

// Dimensions, work list and device buffers of a new problem on every GPU of the handle; the copy streams are ordered after
// whatever the compute streams still have in flight (the previous problem's kernels read the buffers about to be overwritten).
static int setup_problem(fpt_handle* h, int o, int v, int pt_slabs = 0)   // pt_slabs: slabs Pt has room for (0: all o)
{
    if (o < 1 || v < 1) return fail("invalid dimensions o=%d v=%d", o, v);
    Problem P{};
    P.o = o; P.v = v;
    P.vp = padded_v(v);
    P.nt = num_tiles(v);
    P.Kp = roundup(v + o, KGROUP);
    P.G = P.Kp / KGROUP;
    P.nb = num_blocks(P.nt);
    P.dbg_flags = h->dbg_flags;
    P.order = h->item_order;
    P.tw_begin = 0;
    P.tw_count = num_triplets(o);   // a new problem starts with the full triplet list
    P.nitems = P.nb * P.tw_count;
    h->o = o; h->v = v;
    h->tw_begin = 0; h->tw_count = P.tw_count; h->nitems = P.nitems;
    // block descriptor table (positions in (i,j,k) instead of orbital numbers): depends on the tiling of the virtual range only,
    // kept across calls of the same shape (the 6 N_atoms calls of a finite-difference gradient)
    if (h->tab_vp != P.vp) {
        h->tab.assign((size_t)P.nb, BlockTabEntry{});
        for (i64 b = 0; b < P.nb; b++) {
            int A, B, C;
            tetra_decode(b, A, B, C);
            make_block(A, B, C, P.vp, h->tab[b].bd);
            h->tab[b].ngemm = make_gemms(h->tab[b].bd, 0, 1, 2, h->tab[b].gemm);
            make_fast_order(h->tab[b]);
        }
        h->tab_vp = P.vp;
    }
    h->block_cost.resize((size_t)P.nb);
    for (i64 b = 0; b < P.nb; b++) h->block_cost[b] = block_cost(h->tab[b], P.G);
    for (Dev* dp : h->devs) {
        Dev& d = *dp;
        CK(cudaSetDevice(d.dev));
        if (d.Pt.ensure((size_t)(pt_slabs ? pt_slabs : o) * P.vp * P.vp * P.Kp * sizeof(double))) return 1;
        d.pt_slabs = pt_slabs ? pt_slabs : o;
        if (d.Qt.ensure((size_t)o * o * P.G * P.vp * KGROUP * sizeof(double))) return 1;
        if (d.OV2.ensure((size_t)ov2_elems(P) * sizeof(double))) return 1;
        if (d.T1d.ensure((size_t)o * v * sizeof(double))) return 1;
        if (d.fo.ensure((size_t)o * sizeof(double))) return 1;
        if (d.fv.ensure((size_t)v * sizeof(double))) return 1;
        if (d.partials.ensure((size_t)d.n_sm * 4 * sizeof(double))) return 1;
        if (d.counter.ensure(sizeof(unsigned long long))) return 1;
        if (d.out.ensure(sizeof(double))) return 1;
        if (d.prof.ensure((size_t)d.n_sm * NPROF * sizeof(long long))) return 1;
        if (d.blocktab.ensure(h->tab.size() * sizeof(BlockTabEntry))) return 1;
        if (d.tab_vp != P.vp) {
            CK(cudaMemcpyAsync(d.blocktab.p, h->tab.data(), h->tab.size() * sizeof(BlockTabEntry), cudaMemcpyHostToDevice, d.stream));
            d.tab_vp = P.vp;
        }
        d.prob = P;
        d.prob.blocktab = (const BlockTabEntry*)d.blocktab.p;
        d.prob.Pt = d.Pt.d(); d.prob.Qt = d.Qt.d(); d.prob.OV2 = d.OV2.d(); d.prob.T1d = d.T1d.d();
        d.prob.fo = d.fo.d(); d.prob.fv = d.fv.d();
        CK(cudaEventRecord(d.ev_start, d.stream));
        CK(cudaStreamWaitEvent(d.copy, d.ev_start, 0));
        if (dp == h->devs[0]) CK(cudaEventRecord(d.tl[0], d.copy));
    }
    return 0;
}

// Pt's padding (rows x,y >= v and kappa >= v+o) must read as zero.  The prep kernels only ever write real entries, so after one
// memset a buffer stays clean for every later problem of the same shape.
static int pt_zero_padding(Dev& d)
{
    const Problem& P = d.prob;
    if (d.clean_o == P.o && d.clean_v == P.v && d.clean_ptr == d.Pt.p && d.clean_slabs >= d.pt_slabs) return 0;
    CK(cudaMemsetAsync(d.Pt.p, 0, (size_t)d.pt_slabs * P.vp * P.vp * P.Kp * sizeof(double), d.stream));
    return 0;
}
static void upload_begin(fpt_handle* h)
{
    h->loaded = false;
    h->launches = 0;
    h->h2d = 0.0;
    h->stage_host_ms = 0.0;
    for (Dev* d : h->devs) d->clean_o = -1;
}
static int upload_end(fpt_handle* h, bool sync)
{
    for (Dev* dp : h->devs) {
        Dev& d = *dp;
        CK(cudaSetDevice(d.dev));
        if (dp == h->devs[0]) {
            CK(cudaEventRecord(d.tl[1], d.copy));
            CK(cudaEventRecord(d.tl[2], d.stream));
        }
        if (sync) {
            CK(cudaStreamSynchronize(d.copy));
            CK(cudaStreamSynchronize(d.stream));
        }
        d.clean_o = d.prob.o; d.clean_v = d.prob.v; d.clean_ptr = d.Pt.p; d.clean_slabs = d.pt_slabs;
    }
    h->loaded = true;
    return 0;
}

// the parts common to all routes that do not depend on a slice of the occupied range: T1 -> T1d; fo, fv; Pt's zero padding
static int upload_t1_f(fpt_handle* h, const double* T1, const double* fo, const double* fv)
{
    const int o = h->o, v = h->v;
    std::vector<const double*> dT1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sT1; }, T1, (size_t)o * v, dT1)) return 1;
    for (size_t i = 0; i < h->devs.size(); i++) {
        Dev& d = *h->devs[i];
        CK(cudaSetDevice(d.dev));
        const bool dev_in = classify(fo) == PK_DEVICE;
        CK(cudaMemcpyAsync(d.fo.p, fo, o * sizeof(double), cudaMemcpyDefault, d.stream));
        CK(cudaMemcpyAsync(d.fv.p, fv, v * sizeof(double), cudaMemcpyDefault, d.stream));
        if (!dev_in) h->h2d += (o + v) * sizeof(double);
        if (pt_zero_padding(d)) return 1;
        prep_t1<<<grid1d(o * v), 256, 0, d.stream>>>(d.prob, d.T1d.d(), dT1[i]);
        CK(cudaGetLastError());
    }
    h->launches += 1;
    return 0;
}

// ---- symmetry-unique halves ---------------------------------------------------------------------------------------------------
// The reference's algorithms are only consistent for inputs that carry the physical index symmetries:
//     OVVV[i,a,b,c] = OVVV[i,a,c,b],   T2[i,j,a,b] = T2[j,i,b,a],   OVOV[i,a,j,b] = OVOV[j,b,i,a].
// For pageable host inputs only the half the symmetry leaves free crosses PCIe -- b <= c of OVVV, a <= b of T2 and OVOV: contiguous
// prefixes of the slowest index's slabs, so the host threads still stream through memory -- and the device kernels write the mirror
// images (prep_pt_particle_tri, expand_*_tri).  The bytes the host has to touch, which bound the end-to-end time of a multi-GPU
// call, drop from 417 to 215 MB at C4.  Arrays are sampled first; any that does not look symmetric is uploaded in full, as is
// everything after fpt_set_symmetric_inputs(h, 0).
static bool sample_symmetric(const double* A, size_t n0, size_t n1, size_t n2, size_t n3, int kind)
{
    // kind 0: A[i,a,b,c] vs A[i,a,c,b];  1: A[i,j,a,b] vs A[j,i,b,a];  2: A[i,a,j,b] vs A[j,b,i,a]   (first index fastest)
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    auto next = [&s](size_t m) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (size_t)((s >> 33) % m); };
    for (int t = 0; t < 512; t++) {
        const size_t i0 = next(n0), i1 = next(n1), i2 = next(n2), i3 = next(n3);
        const double x = A[i0 + n0 * (i1 + n1 * (i2 + n2 * i3))];
        double y;
        if (kind == 0) y = A[i0 + n0 * (i1 + n1 * (i3 + n2 * i2))];
        else if (kind == 1) y = A[i1 + n0 * (i0 + n1 * (i3 + n2 * i2))];
        else y = A[i2 + n0 * (i3 + n1 * (i0 + n2 * i1))];
        if (fabs(x - y) > 1e-12 * (fabs(x) + fabs(y)) + 1e-300) return false;
    }
    return true;
}
static bool use_half(const fpt_handle* h, const double* A, size_t n0, size_t n1, size_t n2, size_t n3, int kind)
{
    if (!h->sym_inputs || classify(A) != PK_PAGEABLE) return false;
    if (n0 * n1 * n2 * n3 * sizeof(double) < ((size_t)4 << 20)) return false;   // not worth a second kernel
    return sample_symmetric(A, n0, n1, n2, n3, kind);
}

// T2 -> Pt hole part and Qt (hole part of Qt from the whole OOOV on the GPUs, dOOOV; empty on the density-fitted route)
static int upload_t2(fpt_handle* h, const double* T2, const std::vector<const double*>& dOOOV, bool with_pt_hole = true)
{
    const int o = h->o, v = h->v;
    const size_t o2 = (size_t)o * o;
    std::vector<const double*> dT2;
    const bool half = use_half(h, T2, o, o, v, v, 1);
    if (half) {
        View vw;
        for (int b = 0; b < v; b++) vw.add((size_t)b * o2 * v * sizeof(double), 1, o2 * (b + 1) * sizeof(double), o2 * (b + 1) * sizeof(double));
        std::vector<const double*> dTri;
        if (distribute(h, [](Dev& d) -> DevBuf& { return d.sTri; }, T2, vw.total / sizeof(double), dTri, &vw)) return 1;
        dT2.resize(h->devs.size());
        for (size_t i = 0; i < h->devs.size(); i++) {
            Dev& d = *h->devs[i];
            CK(cudaSetDevice(d.dev));
            if (d.sT2.ensure(o2 * v * v * sizeof(double))) return 1;
            expand_t2_tri<<<grid1d((i64)o2 * v * v), 256, 0, d.stream>>>(o, v, d.sT2.d(), dTri[i]);
            CK(cudaGetLastError());
            dT2[i] = d.sT2.d();
        }
        h->launches += 1;
    } else if (distribute(h, [](Dev& d) -> DevBuf& { return d.sT2; }, T2, o2 * v * v, dT2)) return 1;
    for (size_t i = 0; i < h->devs.size(); i++) {
        Dev& d = *h->devs[i];
        CK(cudaSetDevice(d.dev));
        const Problem& P = d.prob;
        d.cur_T2 = dT2[i];
        if (with_pt_hole) prep_pt_hole<<<grid1d((i64)o2 * v * v), 256, 0, d.stream>>>(P, d.Pt.d(), dT2[i], 0, o, nullptr);
        prep_qt<<<grid1d((i64)o2 * P.G * P.vp * KGROUP), 256, 0, d.stream>>>(P, d.Qt.d(), dT2[i], dOOOV.empty() ? nullptr : dOOOV[i]);
        CK(cudaGetLastError());
    }
    h->launches += 2;
    return 0;
}

// OVOV -> OV2
static int upload_ovov(fpt_handle* h, const double* OVOV)
{
    const int o = h->o, v = h->v;
    const size_t o2 = (size_t)o * o;
    std::vector<const double*> dOVOV;
    const bool half = use_half(h, OVOV, o, v, o, v, 2);
    if (half) {
        View vw;   // for every (b, j): the prefix a <= b of the (i, a) plane
        for (int b = 0; b < v; b++)
            vw.add((size_t)b * o2 * v * sizeof(double), (size_t)o, (size_t)o * (b + 1) * sizeof(double), (size_t)o * v * sizeof(double));
        std::vector<const double*> dTri;
        if (distribute(h, [](Dev& d) -> DevBuf& { return d.sTri2; }, OVOV, vw.total / sizeof(double), dTri, &vw)) return 1;
        dOVOV.resize(h->devs.size());
        for (size_t i = 0; i < h->devs.size(); i++) {
            Dev& d = *h->devs[i];
            CK(cudaSetDevice(d.dev));
            if (d.sOVOV.ensure(o2 * v * v * sizeof(double))) return 1;
            expand_ovov_tri<<<grid1d((i64)o2 * v * v), 256, 0, d.stream>>>(o, v, d.sOVOV.d(), dTri[i]);
            CK(cudaGetLastError());
            dOVOV[i] = d.sOVOV.d();
        }
        h->launches += 1;
    } else if (distribute(h, [](Dev& d) -> DevBuf& { return d.sOVOV; }, OVOV, o2 * v * v, dOVOV)) return 1;
    for (size_t i = 0; i < h->devs.size(); i++) {
        Dev& d = *h->devs[i];
        CK(cudaSetDevice(d.dev));
        prep_ov2<<<grid1d(ov2_elems(d.prob)), 256, 0, d.stream>>>(d.prob, d.OV2.d(), dOVOV[i]);
        CK(cudaGetLastError());
    }
    h->launches += 1;
    return 0;
}

// OVVV[p0 : p0+np, :, :, :] -> Pt particle part on every GPU.
//  * phase 0: in chunks of at most 64 MB over the slowest index c; chunk n is staged (and gathered) into buffer n & 1 while the prep
//    kernel of chunk n-1 runs out of the other one;
//  * phase > 0 (later phases of a split call, see triples_conv): one transfer of the whole sub-block into the phase's own buffer --
//    nothing on the compute stream can run before the previous phase's kernel has finished, so a buffer could not be recycled anyway.
//  * half: only the prefix b <= c of every c crosses PCIe, the mirror image is written on the device (see use_half).
// A proper sub-range of p is a view of rows of np doubles, o apart, and arrives packed.
static int upload_ovvv(fpt_handle* h, const double* OVVV, int p0, int np, int phase, bool half)
{
    const bool chunked = phase == 0;
    const int o = h->o, v = h->v, L = (int)h->devs.size();
    const size_t ov = (size_t)o * v, npv = (size_t)np * v;
    const bool whole = (p0 == 0 && np == o);
    const bool on_dev = classify(OVVV) == PK_DEVICE;
    if (on_dev && (!whole || half)) return fail("internal: views of device-resident arrays are not supported");
    const size_t budget = (chunked && (!on_dev || L > 1)) ? (size_t)64 << 20 : ~(size_t)0;
    std::vector<const double*> dChunk;
    int n = 0;
    for (int c0 = 0; c0 < v; n++) {
        // chunk [c0, c0 + cn): as many c as fit the budget (at least one)
        int cn = 0;
        size_t elems = 0;
        while (c0 + cn < v) {
            const size_t add = npv * (half ? (size_t)(c0 + cn + 1) : (size_t)v);
            if (cn > 0 && (elems + add) * sizeof(double) > budget) break;
            elems += add;
            cn++;
        }
        const int bsel = n & 1;
        if (chunked && n >= 2)
            for (Dev* dp : h->devs) {   // the buffer's previous content has been consumed
                CK(cudaSetDevice(dp->dev));
                CK(cudaStreamWaitEvent(dp->copy, dp->ev_free[bsel], 0));
            }
        auto buf = [bsel, phase](Dev& d) -> DevBuf& { return phase == 0 ? d.sChunk[bsel] : d.sPhase[phase]; };
        if (whole && !half) {
            if (distribute(h, buf, OVVV + (size_t)c0 * ov * v, elems, dChunk)) return 1;
        } else {
            View vw;
            for (int c = c0; c < c0 + cn; c++)
                vw.add(((size_t)c * ov * v + p0) * sizeof(double), (size_t)v * (half ? c + 1 : v), (size_t)np * sizeof(double), (size_t)o * sizeof(double));
            if (distribute(h, buf, OVVV, elems, dChunk, &vw)) return 1;
        }
        for (int i = 0; i < L; i++) {
            Dev& d = *h->devs[i];
            CK(cudaSetDevice(d.dev));
            const unsigned gx = (unsigned)((npv + 31) / 32);
            if (half) {
                prep_pt_particle_tri<<<dim3(gx, (unsigned)((cn + 31) / 32), (unsigned)(c0 + cn)), dim3(32, 8), 0, d.stream>>>(d.prob, d.Pt.d(), dChunk[i], c0, cn, p0, np, 0);
                prep_pt_particle_tri<<<dim3(gx, (unsigned)((c0 + cn + 30) / 32 + 1), (unsigned)cn), dim3(32, 8), 0, d.stream>>>(d.prob, d.Pt.d(), dChunk[i], c0, cn, p0, np, 1);
            } else {
                prep_pt_particle<<<dim3(gx, (unsigned)((cn + 31) / 32), (unsigned)v), dim3(32, 8), 0, d.stream>>>(d.prob, d.Pt.d(), dChunk[i], c0, cn, p0, np);
            }
            CK(cudaGetLastError());
            if (chunked) CK(cudaEventRecord(d.ev_free[bsel], d.stream));
        }
        h->launches += half ? 2 : 1;
        c0 += cn;
    }
    return 0;
}

// Conventional upload, OVVV in the occupied slices pb[0] = 0 < pb[1] < ... < pb[nph] = o, one after the other; after slice t,
// `after_slice(t)` may enqueue work on the compute streams (the kernel over the triplets with i < pb[t+1], see triples_conv).
// Everything else goes first, whole.
template <class After>
static int upload_conv_slices(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV, const double* OOOV,
                              const double* OVOV, const double* fo, const double* fv, const int* pb, int nph, After after_slice)
{
    if (setup_problem(h, o, v)) return 1;
    if (upload_t1_f(h, T1, fo, fv)) return 1;
    std::vector<const double*> dOOOV;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sOOOV; }, OOOV, (size_t)o * o * o * v, dOOOV)) return 1;
    if (upload_t2(h, T2, dOOOV)) return 1;
    if (upload_ovov(h, OVOV)) return 1;
    const bool half = use_half(h, OVVV, o, v, v, v, 0);
    for (int t = 0; t < nph; t++) {
        if (upload_ovvv(h, OVVV, pb[t], pb[t + 1] - pb[t], t, half)) return 1;
        if (after_slice(t)) return 1;
    }
    return 0;
}

static int upload_conv_impl(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                            const double* OOOV, const double* OVOV, const double* fo, const double* fv, bool sync)
{
    const int pb[2] = {0, o};
    if (upload_conv_slices(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, pb, 1, [](int) { return 0; })) return 1;
    return upload_end(h, sync);
}

extern "C" int fpt_upload_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                               const double* OOOV, const double* OVOV, const double* fo, const double* fv)
{
    if (check_idle(h, "fpt_upload_conv")) return 1;
    if (!T1 || !T2 || !OVVV || !OOOV || !OVOV || !fo || !fv) return fail("fpt_upload_conv: NULL array argument");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_upload_conv", {T1, T2, OVVV, OOOV, OVOV, fo, fv})) return 1;
    upload_begin(h);
    if (upload_conv_impl(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, true)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.upload_ms = ms_since(t0);
    return 0;
}

// DF route: the (ia|bd), (ij|ka), (ia|jb) blocks that DFERI.jl:88-180 would materialise on the host are assembled on the GPU
// from the B factors, straight into the fused kernel's layouts.  With several GPUs the big one -- Pt's particle part,
// 2 naux o v^3 flops -- is assembled in slices over the occupied index p, one slice per GPU, and the slices are exchanged over
// NVLink (one broadcast per owner, grouped); the two small ones are built redundantly everywhere.
static int upload_df_impl(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                          const double* BOV, const double* BVV, const double* fo, const double* fv, bool sync)
{
    if (setup_problem(h, o, v)) return 1;
    const int L = (int)h->devs.size(), W = h->world;
    std::vector<const double*> dBOO, dBOV, dBVV;
    if (upload_t1_f(h, T1, fo, fv)) return 1;
    // Pt hole part and Qt particle part from T2; Qt's hole part OOOV[l,q,r,z] = sum_Q BOO[Q,l,q] BOV[Q,r,z] below   (DFERI.jl:88-112)
    if (upload_t2(h, T2, {})) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOO; }, BOO, (size_t)naux * o * o, dBOO)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOV; }, BOV, (size_t)naux * o * v, dBOV)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBVV; }, BVV, (size_t)naux * v * v, dBVV)) return 1;
    for (int i = 0; i < L; i++) {
        Dev& d = *h->devs[i];
        CK(cudaSetDevice(d.dev));
        const Problem& P = d.prob;
        GemmOut out{};
        out.P = P;
        out.C = d.Qt.d();
        CK(gemm_tn_launch<EPI_QT_HOLE>(d.stream, dBOO[i], rowmap_identity(), dBOV[i], rowmap_identity(), (i64)o * o, o * v, naux, out));
        // OV2: OVOV[q,y,r,z] = sum_Q BOV[Q,q,y] BOV[Q,r,z]                                            (DFERI.jl:139-154)
        CK(cudaMemsetAsync(d.OV2.p, 0, (size_t)ov2_elems(P) * sizeof(double), d.stream));
        out.C = d.OV2.d();
        CK(gemm_tn_launch<EPI_OV2>(d.stream, dBOV[i], rowmap_identity(), dBOV[i], rowmap_identity(), (i64)o * v, o * v, naux, out));
        // Pt particle part, this GPU's slice of p: OVVV[p,y,x,d] = sum_Q BOV[Q,p,y] BVV[Q,x,d]       (DFERI.jl:156-180)
        const int p0 = (int)((i64)o * d.grank / W), p1 = (int)((i64)o * (d.grank + 1) / W);
        out.C = d.Pt.d();
        out.p0 = p0;
        const RowMap mA{p0, o, 1, v};   // m = y + v*pl  ->  BOV row (p0+pl) + o*y
        const RowMap mB{0, v, 1, v};    // n = d + v*x   ->  BVV row x + v*d
        CK(gemm_tn_launch<EPI_PT>(d.stream, dBOV[i], mA, dBVV[i], mB, (i64)(p1 - p0) * v, v * v, naux, out));
    }
    h->launches += 3;
    if (W > 1) {
        const size_t pslab = (size_t)h->devs[0]->prob.vp * h->devs[0]->prob.vp * h->devs[0]->prob.Kp;
        NCK(nccl_api().GroupStart());
        for (int i = 0; i < L; i++) {
            Dev& d = *h->devs[i];
            for (int g = 0; g < W; g++) {
                const int p0 = (int)((i64)o * g / W), p1 = (int)((i64)o * (g + 1) / W);
                if (p1 > p0) NCK(nccl_api().Broadcast(d.Pt.d() + p0 * pslab, d.Pt.d() + p0 * pslab, (p1 - p0) * pslab, ncclDouble, g, d.comm, d.stream));
            }
        }
        NCK(nccl_api().GroupEnd());
    }
    return upload_end(h, sync);
}

extern "C" int fpt_upload_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                             const double* BOV, const double* BVV, const double* fo, const double* fv)
{
    if (check_idle(h, "fpt_upload_df")) return 1;
    if (!T1 || !T2 || !BOO || !BOV || !BVV || !fo || !fv) return fail("fpt_upload_df: NULL array argument");
    if (naux < 1) return fail("fpt_upload_df: invalid naux=%d", naux);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_upload_df", {T1, T2, BOO, BOV, BVV, fo, fv})) return 1;
    upload_begin(h);
    if (upload_df_impl(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, true)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.upload_ms = ms_since(t0);
    return 0;
}

extern "C" int fpt_num_items(fpt_handle* h, long long* n)
{
    if (!h || !n) return fail("fpt_num_items: NULL argument");
    if (!h->loaded) return fail("fpt_num_items: no problem uploaded");
    *n = h->nitems;
    return 0;
}

// ---- compute ---------------------------------------------------------------------------------------------------------------
static Problem current_problem(const fpt_handle* h, const Dev& d)
{
    Problem P = d.prob;
    P.order = h->item_order;
    P.dbg_flags = h->dbg_flags | (h->deterministic ? 512 : 0);
    P.tw_begin = h->tw_begin;
    P.tw_count = h->tw_count;
    P.nitems = h->nitems;
    return P;
}

// Static split of the item range [b, e) into `world` contiguous parts of equal estimated cost (shard_items in fpt_layout.h,
// shared with the CPU emulator so that the gloo tests exercise the very same split)
static void shard_range(const fpt_handle* h, const Problem& P, i64 b, i64 e, int rank, int world, i64* sb, i64* se)
{
    shard_items(P, h->block_cost.data(), b, e, rank, world, sb, se);
}

// launch the fused kernel + reduction for [item_begin, item_end) of the work list described by P on one GPU (asynchronous; the result
// is stored in d.out, or added to it for the second phase of a split call, which also has its own pair of timing events)
static int compute_launch(fpt_handle* h, Dev& d, const Problem& P, i64 item_begin, i64 item_end, int phase, int accumulate = -1,
                          RingMap ring = RingMap{nullptr, nullptr})
{
    // phase >= 0: the launch is timed with event pair `phase`; accumulate (default: phase > 0): add to d.out instead of storing
    if (accumulate < 0) accumulate = phase > 0;
    CK(cudaSetDevice(d.dev));
    const i64 n = item_end - item_begin;
    int grid = d.n_sm;
    if ((i64)grid > n) grid = (int)(n > 0 ? n : 1);
    CK(cudaMemsetAsync(d.counter.p, 0, sizeof(unsigned long long), d.stream));
    if (phase >= 0) CK(cudaEventRecord(d.ev0[phase], d.stream));
    unsigned long long* ctr = (unsigned long long*)d.counter.p;
#ifdef FPT_WITH_VARIANT2
    if (h->kernel_variant == 2) {
        if (h->profiling) triples_kernel2<true><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p);
        else triples_kernel2<false><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p);
    } else
#endif
    if (ring.trips)   // slab ring of the DF route: explicit triplet list, Pt through the slot map
        triples_kernel<false, true><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    else if (h->profiling)
        triples_kernel<true><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    else
        triples_kernel<false><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    d.last_grid = grid;
    CK(cudaGetLastError());
    if (phase >= 0) CK(cudaEventRecord(d.ev1[phase], d.stream));
    reduce_partials<<<1, 32, 0, d.stream>>>(d.partials.d(), grid, d.out.d(), accumulate);
    CK(cudaGetLastError());
    d.shard_b = item_begin;
    d.shard_e = item_end;
    return 0;
}

// Enqueue the kernels for items [item_begin, item_end) of the work list over the triplet window [tw_begin, tw_begin + tw_count):
// every GPU of the communicator takes its static, cost-weighted shard.
static int compute_launch_all(fpt_handle* h, i64 tw_begin, i64 tw_count, i64 item_begin, i64 item_end, int phase)
{
    for (Dev* dp : h->devs) {
        Problem P = current_problem(h, *dp);
        P.tw_begin = tw_begin;
        P.tw_count = tw_count;
        P.nitems = P.nb * tw_count;
        i64 sb, se;
        shard_range(h, P, item_begin, item_end, dp->grank, h->world, &sb, &se);
        if (compute_launch(h, *dp, P, sb, se, phase)) return 1;
    }
    h->last_profiled = h->profiling;
    return 0;
}

// E(T) is one scalar all-reduce; the 8-byte result is sent to the host.  Nothing here waits for the GPU.
static int compute_collect(fpt_handle* h, i64 n_items)
{
    if (h->world > 1) {
        NCK(nccl_api().GroupStart());
        for (Dev* dp : h->devs) NCK(nccl_api().AllReduce(dp->out.p, dp->out.p, 1, ncclDouble, ncclSum, dp->comm, dp->stream));
        NCK(nccl_api().GroupEnd());
    }
    Dev& d0 = *h->devs[0];
    CK(cudaSetDevice(d0.dev));
    CK(cudaMemcpyAsync(h->res_pinned, d0.out.p, sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    CK(cudaEventRecord(d0.tl[5], d0.stream));
    h->pend_items = n_items;
    return 0;
}

static int compute_enqueue(fpt_handle* h, i64 item_begin, i64 item_end)
{
    if (item_end < 0 || item_end > h->nitems) item_end = h->nitems;
    if (item_begin < 0) item_begin = 0;
    if (item_begin > item_end) item_begin = item_end;
    h->nphase = 1;
    if (compute_launch_all(h, h->tw_begin, h->tw_count, item_begin, item_end, 0)) return 1;
    return compute_collect(h, item_end - item_begin);
}

static int compute_finish(fpt_handle* h, double* Et, fpt_stats* st)
{
    float ms_max = 0.f;
    for (Dev* dp : h->devs) {
        CK(cudaSetDevice(dp->dev));
        CK(cudaStreamSynchronize(dp->stream));
        float sum = 0.f;
        for (int t = 0; t < h->nphase; t++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, dp->ev0[t], dp->ev1[t]));
            sum += ms;
        }
        if (sum > ms_max) ms_max = sum;
    }
    if (Et) *Et = *h->res_pinned;
    // algorithmic flops of the triplets in the window, scaled by the share of the window's items that were computed
    const double ntrip = (double)h->tw_count;
    const int v = h->v, o = h->o;
    h->last.kernel_ms = ms_max;
    h->last.n_items = h->pend_items;
    h->last.n_triplets = (long long)ntrip;
    h->last.flops = 12.0 * v * (double)v * v * (v + o) * ntrip * (h->nitems ? (double)h->pend_items / (double)h->nitems : 0.0);
    h->last.n_launches = h->launches + 2 * h->nphase * (int)h->devs.size();
    h->last.n_sm = h->devs[0]->n_sm;
    // timeline of the first GPU, milliseconds since the upload began (entries stay 0 when the compute followed an older upload)
    Dev& d0 = *h->devs[0];
    CK(cudaSetDevice(d0.dev));
    for (double& t : h->timeline) t = 0.0;
    h->timeline[0] = h->stage_host_ms;
    cudaEvent_t marks[5] = {d0.tl[1], d0.tl[2], d0.ev0[0], d0.ev1[h->nphase - 1], d0.tl[5]};
    for (int t = 0; t < 5; t++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d0.tl[0], marks[t]) == cudaSuccess) h->timeline[1 + t] = ms;
        else cudaGetLastError();
    }
    if (st) *st = h->last;
    return 0;
}

extern "C" int fpt_compute(fpt_handle* h, long long item_begin, long long item_end, double* Et, fpt_stats* st)
{
    if (!h || !Et) return fail("fpt_compute: NULL argument");
    if (check_idle(h, "fpt_compute")) return 1;
    if (!h->loaded) return fail("fpt_compute: no problem uploaded");
    DeviceGuard guard;
    if (compute_enqueue(h, item_begin, item_end)) return 1;
    return compute_finish(h, Et, st);
}

// Milliseconds since the start of the last upload, on the first GPU's clock: out8 = {host time spent copying pageable memory into
// the pinned ring (wall, overlaps the DMAs), last H2D done, operands ready (gathers + prep done), kernel begin, kernel end,
// result on its way to the host, 0, 0}.
extern "C" int fpt_last_timeline(fpt_handle* h, double* out8)
{
    if (!h || !out8) return fail("fpt_last_timeline: NULL argument");
    for (int t = 0; t < 8; t++) out8[t] = h->timeline[t];
    return 0;
}

// ---- one-call forms ----------------------------------------------------------------------------------------------------------
// upload and compute are enqueued back to back (no host synchronisation in between); `async` returns as soon as the caller's
// arrays have been consumed, fpt_wait collects the result.
static int finish_tail(fpt_handle* h, bool async, wall::time_point t0, double* Et, fpt_stats* st)
{
    if (async) {
        for (Dev* dp : h->devs) {   // inputs in pinned memory are read by the DMA engines directly: wait for those reads
            CK(cudaSetDevice(dp->dev));
            CK(cudaStreamSynchronize(dp->copy));
        }
        h->last.upload_ms = ms_since(t0);
        h->pending = true;
        return 0;
    }
    if (compute_finish(h, Et, nullptr)) return 1;
    h->last.total_ms = ms_since(t0);
    h->last.upload_ms = h->timeline[2];
    if (st) *st = h->last;
    return 0;
}
static int finish_call(fpt_handle* h, bool async, wall::time_point t0, double* Et, fpt_stats* st)
{
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    if (compute_enqueue(h, 0, -1)) return 1;
    return finish_tail(h, async, t0, Et, st);
}

// Phases of a one-call conventional evaluation.  The triplets with i < pb only read the operands of the occupied indices p < pb, and
// they are the first num_triplets(pb) entries of the reference's triplet list (ijk.jl:49,63,83 loops i slowest).  So the call is cut
// at occupied boundaries 0 = pb[0] < pb[1] < ... < pb[n] = o:  upload everything but OVVV, then OVVV[p < pb[1]]; launch the kernel over
// that window; while it runs, the host threads and the DMA engines bring OVVV[pb[1] <= p < pb[2]]; and so on.  The kernel starts after
// 1/n of OVVV has arrived, and the rest of the host-bound staging time -- the part of an 8-GPU call that does not shrink with the
// number of GPUs -- disappears behind the kernels: phase t holds (pb[t+1]^3 - pb[t]^3) / o^3 of the work, which covers the staging
// of slice t+1 as long as the whole kernel takes longer than the whole upload.
// Conditions: host-resident OVVV small enough that the later slices (packed copies on every GPU) are cheap to hold, enough occupied
// orbitals, a full default work list.  Boundaries are multiples of 4 (32-byte rows for the streaming copies).
// FERMI_PT_B200_SPLIT = number of phases wanted (default 2; 0 or 1: no split; more phases start the kernel earlier but re-read
// more of the host array's cache lines -- a slice is 8 (o / nph) bytes of every 8 o-byte row; measured at C4: no gain beyond 2).
static int split_points(fpt_handle* h, int o, int v, const double* T2, const double* OVVV, const double* OVOV, int* pb)
{
    int want = 2;
    if (const char* e = getenv("FERMI_PT_B200_SPLIT")) want = atoi(e);
    if (want > MAX_PHASES) want = MAX_PHASES;
    pb[0] = 0;
    pb[1] = o;
    if (want < 2 || o < 8 || classify(OVVV) == PK_DEVICE) return 1;
    if ((double)o * v * v * v * sizeof(double) > 4e9) return 1;
    if (h->dbg_flags || h->profiling || h->item_order != 1) return 1;
    int n = 0;
    for (int t = 1; t < want; t++) {
        const int b = (int)(((i64)o * t / want + 2) & ~3);
        if (b > pb[n] && b < o) pb[++n] = b;
    }
    pb[++n] = o;
    return n;
}

static int triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV, const double* OOOV,
                        const double* OVOV, const double* fo, const double* fv, double* Et, fpt_stats* st, bool async, const char* who)
{
    if (check_idle(h, who)) return 1;
    if (!T1 || !T2 || !OVVV || !OOOV || !OVOV || !fo || !fv || (!async && !Et)) return fail("%s: NULL argument", who);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, who, {T1, T2, OVVV, OOOV, OVOV, fo, fv})) return 1;
    upload_begin(h);
    int pb[MAX_PHASES + 1];
    const int nph = split_points(h, o, v, T2, OVVV, OVOV, pb);
    if (nph <= 1) {
        if (upload_conv_impl(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, false)) return 1;
        return finish_call(h, async, t0, Et, st);
    }
    h->nphase = nph;
    auto launch_phase = [&](int t) -> int {
        if (t == nph - 1 && upload_end(h, false)) return 1;
        const i64 nb = h->devs[0]->prob.nb;
        const i64 u0 = num_triplets(pb[t]), u1 = num_triplets(pb[t + 1]);
        return compute_launch_all(h, u0, u1 - u0, 0, nb * (u1 - u0), t);
    };
    if (upload_conv_slices(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, pb, nph, launch_phase)) return 1;
    const i64 nb = h->devs[0]->prob.nb;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    if (compute_collect(h, nb * num_triplets(o))) return 1;
    return finish_tail(h, async, t0, Et, st);
}

// ---- DF route without the full (ov|vv) block: a ring of occupied slabs (north star: "assembling (bd|ai) slices on the fly") ---------
// Pt holds o slabs of vp^2 Kp doubles (22.9 GB at C5; o = 100, v = 800 would need 467 GB).  Every slab is a GEMM away from the B
// factors (2 naux v^3 flops), so Pt need not be resident: the occupied range is cut into blocks of `ob`, the triplet list is walked
// block triple by block triple (I >= J >= K; i in I, j in J, k in K), and the ring holds just the 3 ob slabs of the current block
// triple -- slot group 0 for I, 1 for J (or I's when J = I), 2 for K (or J's when K = J).  Going to the next K assembles ob slabs;
// a block triple of distinct blocks carries ob^3 triplets of 12 v^3 (v + o) flops each, so the re-assembly costs
// naux / (6 ob^2 (v + o)) of the work: 4 % at C3 with ob = 4, 65 % with ob = 1 (three slabs in all).  Per block triple: one
// assembly launch group + one launch of the fused kernel over its explicit triplet list (Problem::trips) with the slot map of that
// block triple (Problem::pslot); E(T) accumulates on the device.  Several GPUs: everyone assembles the same slabs and takes its
// cost-weighted shard of every launch.
struct RingPhase { int I, J, K; i64 trip_off, ntrip; };

static int assemble_slabs(fpt_handle* h, Dev& d, const Problem& P, const int* pslot, int naux, const double* dBOV, const double* dBVV, int p0, int p1)
{
    if (p1 <= p0) return 0;
    const int o = P.o, v = P.v;
    prep_pt_hole<<<grid1d((i64)(p1 - p0) * o * v * v), 256, 0, d.stream>>>(P, d.Pt.d(), d.cur_T2, p0, p1 - p0, pslot);
    CK(cudaGetLastError());
    GemmOut out{};
    out.P = P;
    out.C = d.Pt.d();
    out.p0 = p0;
    out.pslot = pslot;
    const RowMap mA{p0, o, 1, v};   // m = y + v*pl  ->  BOV row (p0+pl) + o*y
    const RowMap mB{0, v, 1, v};    // n = d + v*x   ->  BVV row x + v*d
    CK(gemm_tn_launch<EPI_PT>(d.stream, dBOV, mA, dBVV, mB, (i64)(p1 - p0) * v, v * v, naux, out, d.n_sm));
    h->launches += 2;
    return 0;
}

static int triples_df_ring(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO, const double* BOV,
                           const double* BVV, const double* fo, const double* fv, int ob)
{
    const int nblk = (o + ob - 1) / ob, L = (int)h->devs.size();
    if (setup_problem(h, o, v, 3 * ob)) return 1;
    std::vector<const double*> dBOO, dBOV, dBVV;
    if (upload_t1_f(h, T1, fo, fv)) return 1;
    if (upload_t2(h, T2, {}, false)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOO; }, BOO, (size_t)naux * o * o, dBOO)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOV; }, BOV, (size_t)naux * o * v, dBOV)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBVV; }, BVV, (size_t)naux * v * v, dBVV)) return 1;
    // the phases: block triples in the order I, J <= I, K <= J; triplet lists in the reference's loop order inside each
    std::vector<RingPhase> phases;
    std::vector<int> trips, pslots;
    auto blk_lo = [&](int B) { return B * ob; };
    auto blk_hi = [&](int B) { return std::min(o, (B + 1) * ob); };
    for (int I = 0; I < nblk; I++)
        for (int J = 0; J <= I; J++)
            for (int K = 0; K <= J; K++) {
                RingPhase ph{I, J, K, (i64)trips.size() / 3, 0};
                for (int i = blk_lo(I); i < blk_hi(I); i++)
                    for (int j = blk_lo(J); j < blk_hi(J) && j <= i; j++)
                        for (int k = blk_lo(K); k < blk_hi(K) && k <= j; k++)
                            if (!(i == j && j == k)) { trips.push_back(i); trips.push_back(j); trips.push_back(k); ph.ntrip++; }
                if (ph.ntrip == 0) continue;
                // slot map of this block triple: I -> group 0, J -> group 1 unless J == I, K -> group 2 unless K == J (or I)
                std::vector<int> m((size_t)o, 0);
                const int gJ = (J == I) ? 0 : 1, gK = (K == J) ? gJ : 2;
                for (int p = blk_lo(I); p < blk_hi(I); p++) m[p] = 0 * ob + (p - blk_lo(I));
                for (int p = blk_lo(J); p < blk_hi(J); p++) m[p] = gJ * ob + (p - blk_lo(J));
                for (int p = blk_lo(K); p < blk_hi(K); p++) m[p] = gK * ob + (p - blk_lo(K));
                pslots.insert(pslots.end(), m.begin(), m.end());
                phases.push_back(ph);
            }
    i64 total = 0;
    for (const RingPhase& ph : phases) total += ph.ntrip;
    if (total != num_triplets(o)) return fail("internal: ring phases hold %lld triplets, expected %lld", (long long)total, (long long)num_triplets(o));
    for (int g = 0; g < L; g++) {
        Dev& d = *h->devs[g];
        CK(cudaSetDevice(d.dev));
        const Problem& P0 = d.prob;
        // Qt hole part and OV2 as on the materialised route                                       (DFERI.jl:88-112, 139-154)
        GemmOut out{};
        out.P = P0;
        out.C = d.Qt.d();
        CK(gemm_tn_launch<EPI_QT_HOLE>(d.stream, dBOO[g], rowmap_identity(), dBOV[g], rowmap_identity(), (i64)o * o, o * v, naux, out, d.n_sm));
        CK(cudaMemsetAsync(d.OV2.p, 0, (size_t)ov2_elems(P0) * sizeof(double), d.stream));
        out.C = d.OV2.d();
        CK(gemm_tn_launch<EPI_OV2>(d.stream, dBOV[g], rowmap_identity(), dBOV[g], rowmap_identity(), (i64)o * v, o * v, naux, out, d.n_sm));
        if (d.ringtab.ensure((trips.size() + pslots.size()) * sizeof(int))) return 1;
        int* dtr = (int*)d.ringtab.p;
        int* dps = dtr + trips.size();
        // pageable sources: the copies are staged by the driver before the call returns, the vectors may go out of scope
        CK(cudaMemcpyAsync(dtr, trips.data(), trips.size() * sizeof(int), cudaMemcpyHostToDevice, d.stream));
        CK(cudaMemcpyAsync(dps, pslots.data(), pslots.size() * sizeof(int), cudaMemcpyHostToDevice, d.stream));
        if (g == 0) { CK(cudaEventRecord(d.tl[1], d.copy)); CK(cudaEventRecord(d.tl[2], d.stream)); }
        CK(cudaEventRecord(d.ev0[0], d.stream));
        int curI = -1, curJ = -1, curK = -1;
        for (size_t t = 0; t < phases.size(); t++) {
            const RingPhase& ph = phases[t];
            Problem P = current_problem(h, d);
            const RingMap ring{dps + t * (size_t)o, dtr + 3 * ph.trip_off};
            P.tw_begin = 0;
            P.tw_count = ph.ntrip;
            P.nitems = P.nb * ph.ntrip;
            if (ph.I != curI) { if (assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.I), blk_hi(ph.I))) return 1; curI = ph.I; curJ = curK = -1; }
            if (ph.J != curJ) { if (ph.J != ph.I && assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.J), blk_hi(ph.J))) return 1; curJ = ph.J; curK = -1; }
            if (ph.K != curK) { if (ph.K != ph.J && assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.K), blk_hi(ph.K))) return 1; curK = ph.K; }
            i64 sb, se;
            shard_range(h, P, 0, P.nitems, d.grank, h->world, &sb, &se);
            if (compute_launch(h, d, P, sb, se, -1, t > 0, ring)) return 1;
        }
        CK(cudaEventRecord(d.ev1[0], d.stream));
        d.clean_o = -1;   // the ring's slabs are not the layout a later materialised upload expects to find clean
    }
    h->launches += 2 * (int)phases.size() - 2;   // compute_finish counts one kernel + one reduction per GPU itself
    h->nphase = 1;
    h->last_profiled = false;
    return 0;
}

// Block size of the slab ring for this problem (0: materialise all o slabs).  fpt_set_df_ring: -1 never, 0 automatic -- ring
// with blocks of 4 when the full Pt would take more than 40 % of the device's memory --, n >= 1 ring with blocks of n.
static int df_ring_block(fpt_handle* h, int o, int v)
{
    if (h->dbg_flags || h->profiling || h->item_order != 1) return 0;
    if (h->df_ring > 0) return std::min(h->df_ring, o);
    if (h->df_ring < 0) return 0;
    size_t free_b = 0, total_b = 0;
    if (cudaSetDevice(h->devs[0]->dev) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
    const double vp = padded_v(v), Kp = roundup(v + o, KGROUP);
    const double full = (double)o * vp * vp * Kp * sizeof(double);
    return (full > 0.4 * (double)total_b && o > 12) ? 4 : 0;
}

static int triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO, const double* BOV,
                      const double* BVV, const double* fo, const double* fv, double* Et, fpt_stats* st, bool async, const char* who)
{
    if (check_idle(h, who)) return 1;
    if (!T1 || !T2 || !BOO || !BOV || !BVV || !fo || !fv || (!async && !Et)) return fail("%s: NULL argument", who);
    if (naux < 1) return fail("%s: invalid naux=%d", who, naux);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, who, {T1, T2, BOO, BOV, BVV, fo, fv})) return 1;
    upload_begin(h);
    const int ob = df_ring_block(h, o, v);
    if (ob > 0) {
        if (triples_df_ring(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, ob)) return 1;
        h->last = fpt_stats{};
        h->last.h2d_bytes = h->h2d;
        if (compute_collect(h, h->nitems)) return 1;
        return finish_tail(h, async, t0, Et, st);
    }
    if (upload_df_impl(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, false)) return 1;
    return finish_call(h, async, t0, Et, st);
}

extern "C" int fpt_triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                                const double* OOOV, const double* OVOV, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    return triples_conv(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, Et, st, false, "fpt_triples_conv");
}
extern "C" int fpt_triples_conv_async(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                                      const double* OOOV, const double* OVOV, const double* fo, const double* fv)
{
    return triples_conv(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, nullptr, nullptr, true, "fpt_triples_conv_async");
}
extern "C" int fpt_triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                              const double* BOV, const double* BVV, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    return triples_df(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, Et, st, false, "fpt_triples_df");
}
extern "C" int fpt_triples_df_async(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                                    const double* BOV, const double* BVV, const double* fo, const double* fv)
{
    return triples_df(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, nullptr, nullptr, true, "fpt_triples_df_async");
}

extern "C" int fpt_wait(fpt_handle* h, double* Et, fpt_stats* st)
{
    if (!h || !Et) return fail("fpt_wait: NULL argument");
    if (!h->pending) return fail("fpt_wait: no asynchronous call is in flight");
    DeviceGuard guard;
    h->pending = false;
    const auto t0 = wall::now();
    if (compute_finish(h, Et, nullptr)) return 1;
    h->last.total_ms = h->last.upload_ms + ms_since(t0);   // host time inside the two calls
    if (st) *st = h->last;
    return 0;
}

// ---- single-precision callers -------------------------------------------------------------------------------
// `@set precision single` makes every array of the reference Float32 (IntegralHelper.jl:58-68).  The f32 entry points take those
// arrays as they are: they cross PCIe in 4-byte form (half the bytes of the Float64 call), are widened on the handle's first GPU
// (widen_f32_kernel) and then take the device-input route of the Float64 call -- the arithmetic is FP64 throughout, so the result is
// the exact (T) energy of the rounded inputs, which the reference's Float32 loops only approximate.
static int widen_inputs(fpt_handle* h, const char* who, std::initializer_list<std::pair<const float*, size_t>> arrays, const double** out)
{
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    CK(cudaEventRecord(d.ev_start, d.stream));
    CK(cudaStreamWaitEvent(d.copy, d.ev_start, 0));
    std::vector<View> views;
    views.reserve(arrays.size());
    std::vector<StagePool::Job> jobs;
    int k = 0;
    for (const auto& a : arrays) {
        if (k >= Dev::NF32) return fail("internal: too many arrays for %s", who);
        if (classify(a.first) == PK_DEVICE) return fail("%s: Float32 inputs must be host memory", who);
        if (d.f32in[k].ensure(a.second * sizeof(float)) || d.f32wide[k].ensure(a.second * sizeof(double))) return 1;
        views.push_back(View::contiguous(a.second * sizeof(float)));
        if (stage_to(h, d, d.f32in[k].p, a.first, views.back(), 0, a.second * sizeof(float), classify(a.first), jobs)) return 1;
        k++;
    }
    if (stage_flush(h, jobs)) return 1;
    if (copy_then_stream(d)) return 1;
    k = 0;
    for (const auto& a : arrays) {
        widen_f32_kernel<<<grid1d((i64)a.second), 256, 0, d.stream>>>(d.f32wide[k].d(), (const float*)d.f32in[k].p, (i64)a.second);
        out[k] = d.f32wide[k].d();
        k++;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d.stream));   // the Float64 call that follows synchronises the device anyway (device-resident inputs)
    return 0;
}

extern "C" int fpt_triples_conv_f32(fpt_handle* h, int o, int v, const float* T1, const float* T2, const float* OVVV, const float* OOOV,
                                    const float* OVOV, const float* fo, const float* fv, double* Et, fpt_stats* st)
{
    if (check_idle(h, "fpt_triples_conv_f32")) return 1;
    if (!T1 || !T2 || !OVVV || !OOOV || !OVOV || !fo || !fv || !Et) return fail("fpt_triples_conv_f32: NULL argument");
    if (o < 1 || v < 1) return fail("invalid dimensions o=%d v=%d", o, v);
    DeviceGuard guard;
    const auto t0 = wall::now();
    const size_t so = o, sv = v;
    const double* w[7];
    const double h2d0 = 0.0;
    upload_begin(h);
    if (widen_inputs(h, "fpt_triples_conv_f32", {{T1, so * sv}, {T2, so * so * sv * sv}, {OVVV, so * sv * sv * sv}, {OOOV, so * so * so * sv},
                                                  {OVOV, so * sv * so * sv}, {fo, so}, {fv, sv}}, w)) return 1;
    const double moved = h->h2d - h2d0;
    const int rc = triples_conv(h, o, v, w[0], w[1], w[2], w[3], w[4], w[5], w[6], Et, st, false, "fpt_triples_conv_f32");
    if (!rc) {
        h->last.h2d_bytes = moved;
        h->last.total_ms = ms_since(t0);
        if (st) *st = h->last;
    }
    return rc;
}

extern "C" int fpt_triples_df_f32(fpt_handle* h, int o, int v, int naux, const float* T1, const float* T2, const float* BOO, const float* BOV,
                                  const float* BVV, const float* fo, const float* fv, double* Et, fpt_stats* st)
{
    if (check_idle(h, "fpt_triples_df_f32")) return 1;
    if (!T1 || !T2 || !BOO || !BOV || !BVV || !fo || !fv || !Et) return fail("fpt_triples_df_f32: NULL argument");
    if (o < 1 || v < 1 || naux < 1) return fail("invalid dimensions o=%d v=%d naux=%d", o, v, naux);
    DeviceGuard guard;
    const auto t0 = wall::now();
    const size_t so = o, sv = v, sq = naux;
    const double* w[7];
    upload_begin(h);
    if (widen_inputs(h, "fpt_triples_df_f32", {{T1, so * sv}, {T2, so * so * sv * sv}, {BOO, sq * so * so}, {BOV, sq * so * sv}, {BVV, sq * sv * sv},
                                                {fo, so}, {fv, sv}}, w)) return 1;
    const double moved = h->h2d;
    const int rc = triples_df(h, o, v, naux, w[0], w[1], w[2], w[3], w[4], w[5], w[6], Et, st, false, "fpt_triples_df_f32");
    if (!rc) {
        h->last.h2d_bytes = moved;
        h->last.total_ms = ms_since(t0);
        if (st) *st = h->last;
    }
    return rc;
}

// ---- DF-CCSD particle-particle ladder ------------------------------------------------------------------------
// Replaces cc_update_T2_v4_term!(newT2, T1, T2, moints::IntegralHelper{T,<:AbstractDFERI}, ::RCCSDa), RCCSDHelper.jl:204-220:
//     tau[i,j,c,d] = T2[i,j,c,d] + T1[i,c] T1[j,d];   for every a:  X_a[c,d,b] = sum_Q BVV[Q,c,a] BVV[Q,d,b];  newT2[:,:,a,:] += tau . X_a
// The reference builds one v^3 slab per a on the CPU and contracts it at once; here the slabs of a group of a (as many as fit 1 GB)
// are assembled by ONE launch of the TN GEMM (M = v n_a, N = v^2, K = naux, written straight into the layout the second GEMM
// reads) and contracted by a second launch (M = o^2, N = v n_a, K = v^2, accumulating into newT2 on the device): the (vv|vv) block
// -- 1.35 GB at C4, 205 GB at C5 -- never exists, and both products run on the FP64 tensor pipe.  2 v^4 (naux + o^2) flops.
// A handle over several GPUs of one process splits the range of a; rank handles (one process per GPU) compute it redundantly.
extern "C" int fpt_ccsd_ladder_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BVV,
                                  double* newT2, fpt_stats* st)
{
    if (check_idle(h, "fpt_ccsd_ladder_df")) return 1;
    if (!T1 || !T2 || !BVV || !newT2) return fail("fpt_ccsd_ladder_df: NULL argument");
    if (o < 1 || v < 1 || naux < 1) return fail("fpt_ccsd_ladder_df: invalid dimensions o=%d v=%d naux=%d", o, v, naux);
    if (classify(newT2) == PK_DEVICE) return fail("fpt_ccsd_ladder_df: newT2 must be host memory");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_ccsd_ladder_df", {T1, T2, BVV})) return 1;
    upload_begin(h);
    h->o = o; h->v = v;
    const size_t o2 = (size_t)o * o, v2 = (size_t)v * v, v3 = v2 * v;
    const int L = (int)h->devs.size(), W = h->rank_mode ? 1 : L;
    for (Dev* dp : h->devs) {   // the copy streams continue from whatever the compute streams still have in flight
        CK(cudaSetDevice(dp->dev));
        CK(cudaEventRecord(dp->ev_start, dp->stream));
        CK(cudaStreamWaitEvent(dp->copy, dp->ev_start, 0));
    }
    std::vector<const double*> dT1, dT2, dBVV, dNew;
    const int world_saved = h->world;
    if (h->rank_mode) h->world = 1;   // every rank works alone here: no sharded upload
    int rc = distribute(h, [](Dev& d) -> DevBuf& { return d.sT1; }, T1, o2 ? (size_t)o * v : 0, dT1) ||
             distribute(h, [](Dev& d) -> DevBuf& { return d.sT2; }, T2, o2 * v2, dT2) ||
             distribute(h, [](Dev& d) -> DevBuf& { return d.sBVV; }, BVV, (size_t)naux * v2, dBVV) ||
             distribute(h, [](Dev& d) -> DevBuf& { return d.xNew; }, newT2, o2 * v2, dNew);
    h->world = world_saved;
    if (rc) return 1;
    double flops = 0.0;
    for (int g = 0; g < L; g++) {
        Dev& d = *h->devs[g];
        CK(cudaSetDevice(d.dev));
        const int a_begin = (int)((i64)v * (h->rank_mode ? 0 : g) / W), a_end = (int)((i64)v * (h->rank_mode ? 1 : g + 1) / W);
        if (d.xTau.ensure(o2 * v2 * sizeof(double))) return 1;
        int na_max = (int)std::max<size_t>(1, ((size_t)1 << 30) / (v3 * sizeof(double)));
        if (na_max > a_end - a_begin) na_max = std::max(1, a_end - a_begin);
        if (d.xSlab.ensure(v3 * na_max * sizeof(double))) return 1;
        double* dnew = (double*)dNew[g];   // on a host-input call this is d.xNew
        if (classify(newT2) != PK_DEVICE && dnew != d.xNew.d()) return fail("internal: newT2 staging buffer");
        CK(cudaEventRecord(d.ev0[0], d.stream));
        ladder_tau_kernel<<<dim3((unsigned)((o2 + 31) / 32), (unsigned)((v2 + 31) / 32)), dim3(32, 8), 0, d.stream>>>(o, v, d.xTau.d(), dT1[g], dT2[g]);
        CK(cudaGetLastError());
        for (int a0 = a_begin; a0 < a_end; a0 += na_max) {
            const int na = std::min(na_max, a_end - a0);
            GemmOut out{};
            out.lv = v; out.lo2 = (int)o2; out.la0 = a0;
            out.C = d.xSlab.d();
            CK(gemm_tn_launch<EPI_LADDER_SLAB>(d.stream, dBVV[g], RowMap{(i64)a0 * v, 1, 0, 0x7fffffff}, dBVV[g], rowmap_identity(), (i64)v * na, (int)v2, naux, out, d.n_sm));
            out.C = dnew;
            CK(gemm_tn_launch<EPI_LADDER_OUT>(d.stream, d.xTau.d(), rowmap_identity(), d.xSlab.d(), rowmap_identity(), (i64)o2, v * na, (int)v2, out, d.n_sm));
            h->launches += 2;
            flops += 2.0 * (double)v2 * v * na * ((double)naux + (double)o2);
        }
        CK(cudaEventRecord(d.ev1[0], d.stream));
        // this GPU's strips newT2[:, :, a_begin:a_end, :] go back to the caller's array
        if (a_end > a_begin)
            CK(cudaMemcpy2DAsync(newT2 + o2 * a_begin, o2 * v * sizeof(double), dnew + o2 * a_begin, o2 * v * sizeof(double),
                                 o2 * (a_end - a_begin) * sizeof(double), v, cudaMemcpyDeviceToHost, d.stream));
    }
    float ms_max = 0.f;
    for (Dev* dp : h->devs) {
        CK(cudaSetDevice(dp->dev));
        CK(cudaStreamSynchronize(dp->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, dp->ev0[0], dp->ev1[0]));
        ms_max = std::max(ms_max, ms);
    }
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.kernel_ms = ms_max;
    h->last.total_ms = ms_since(t0);
    h->last.flops = flops;
    h->last.n_launches = h->launches + L;
    h->last.n_sm = h->devs[0]->n_sm;
    if (st) *st = h->last;
    return 0;
}

// ---- MP2 energy ----------------------------------------------------------------------------------------------
// Replaces RMP2_energy(ints::IntegralHelper{T,<:AbstractDFERI,RHFOrbitals}, alg), RMP2a.jl:91-143 (per pair i <= j: B_i^T B_j, then the
// (a,b) sum), and its conventional twin RMP2a.jl:146-169.  DF: (ia|jb) = sum_Q BOV[Q,i,a] BOV[Q,j,b] is ONE launch of the TN GEMM
// (M = N = o v, K = naux) into the reference's OVOV layout; the energy is a fixed-order reduction over it (mp2_energy_kernel).
// Runs on the handle's first GPU (rank handles: every rank computes it).
static int mp2_from_ovov(fpt_handle* h, Dev& d, int o, int v, const double* dOVOV, const double* fo, const double* fv, double flops,
                         wall::time_point t0, double* Emp2, fpt_stats* st)
{
    CK(cudaMemcpyAsync(d.fo.p, fo, o * sizeof(double), cudaMemcpyDefault, d.stream));
    CK(cudaMemcpyAsync(d.fv.p, fv, v * sizeof(double), cudaMemcpyDefault, d.stream));
    const int nblk = d.n_sm * 4;
    mp2_energy_kernel<<<nblk, 256, 0, d.stream>>>(o, v, dOVOV, d.fo.d(), d.fv.d(), d.partials.d());
    reduce_partials<<<1, 32, 0, d.stream>>>(d.partials.d(), nblk, d.out.d(), 0);
    CK(cudaGetLastError());
    CK(cudaEventRecord(d.ev1[0], d.stream));
    CK(cudaMemcpyAsync(h->res_pinned, d.out.p, sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    CK(cudaStreamSynchronize(d.stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, d.ev0[0], d.ev1[0]));
    *Emp2 = *h->res_pinned;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.kernel_ms = ms;
    h->last.total_ms = ms_since(t0);
    h->last.flops = flops;
    h->last.n_launches = h->launches + 2;
    h->last.n_sm = d.n_sm;
    if (st) *st = h->last;
    return 0;
}
static int mp2_begin(fpt_handle* h, Dev& d, int o, int v)
{
    CK(cudaSetDevice(d.dev));
    upload_begin(h);
    h->o = o; h->v = v;
    if (d.fo.ensure(o * sizeof(double)) || d.fv.ensure(v * sizeof(double)) || d.partials.ensure((size_t)d.n_sm * 4 * sizeof(double)) ||
        d.out.ensure(sizeof(double)))
        return 1;
    CK(cudaEventRecord(d.ev_start, d.stream));
    CK(cudaStreamWaitEvent(d.copy, d.ev_start, 0));
    return 0;
}
// one array to the first GPU only (MP2 runs there)
template <class BufOf>
static int to_first_gpu(fpt_handle* h, BufOf bufof, const double* src, size_t n, const double** out)
{
    std::vector<Dev*> all = h->devs;
    const int world = h->world;
    h->devs.resize(1);
    h->world = 1;
    std::vector<const double*> o1;
    const int rc = distribute(h, bufof, src, n, o1);
    h->devs = all;
    h->world = world;
    if (!rc) *out = o1[0];
    return rc;
}

extern "C" int fpt_mp2_df(fpt_handle* h, int o, int v, int naux, const double* BOV, const double* fo, const double* fv, double* Emp2,
                          fpt_stats* st)
{
    if (check_idle(h, "fpt_mp2_df")) return 1;
    if (!BOV || !fo || !fv || !Emp2) return fail("fpt_mp2_df: NULL argument");
    if (o < 1 || v < 1 || naux < 1) return fail("fpt_mp2_df: invalid dimensions o=%d v=%d naux=%d", o, v, naux);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_mp2_df", {BOV, fo, fv})) return 1;
    Dev& d = *h->devs[0];
    if (mp2_begin(h, d, o, v)) return 1;
    const i64 ov = (i64)o * v;
    const double* dBOV = nullptr;
    if (to_first_gpu(h, [](Dev& dd) -> DevBuf& { return dd.sBOV; }, BOV, (size_t)naux * ov, &dBOV)) return 1;
    if (d.xOVOV.ensure((size_t)ov * ov * sizeof(double))) return 1;
    CK(cudaEventRecord(d.ev0[0], d.stream));
    GemmOut out{};
    out.C = d.xOVOV.d();
    out.ldc = ov;
    CK(gemm_tn_launch<EPI_COLMAJOR>(d.stream, dBOV, rowmap_identity(), dBOV, rowmap_identity(), ov, (int)ov, naux, out, d.n_sm));
    h->launches += 1;
    return mp2_from_ovov(h, d, o, v, d.xOVOV.d(), fo, fv, 2.0 * (double)ov * ov * naux, t0, Emp2, st);
}

extern "C" int fpt_mp2_conv(fpt_handle* h, int o, int v, const double* OVOV, const double* fo, const double* fv, double* Emp2, fpt_stats* st)
{
    if (check_idle(h, "fpt_mp2_conv")) return 1;
    if (!OVOV || !fo || !fv || !Emp2) return fail("fpt_mp2_conv: NULL argument");
    if (o < 1 || v < 1) return fail("fpt_mp2_conv: invalid dimensions o=%d v=%d", o, v);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_mp2_conv", {OVOV, fo, fv})) return 1;
    Dev& d = *h->devs[0];
    if (mp2_begin(h, d, o, v)) return 1;
    const double* dOVOV = nullptr;
    if (to_first_gpu(h, [](Dev& dd) -> DevBuf& { return dd.sOVOV; }, OVOV, (size_t)o * v * o * v, &dOVOV)) return 1;
    CK(cudaEventRecord(d.ev0[0], d.stream));
    return mp2_from_ovov(h, d, o, v, dOVOV, fo, fv, 0.0, t0, Emp2, st);
}

// ---- AO -> MO route (replaces Chonky.jl:28-114 for the three blocks the (T) path reads) --------------------------
// C[m + ldc*n] = sum_q A[q + Q*m] B[q + Q*n]
static int quarter(fpt_handle* h, Dev& d, double* C, const double* A, const double* B, i64 M, int N, int Q, i64 ldc = 0)
{
    GemmOut out{};
    out.C = C;
    out.ldc = ldc ? ldc : M;
    CK(gemm_tn_launch<EPI_COLMAJOR>(d.stream, A, rowmap_identity(), B, rowmap_identity(), M, N, Q, out));
    h->launches += 1;
    return 0;
}

// AOERI[mu,nu,rho,sigma] (nbf^4, column-major, chemist notation as in aoints["ERI"]), Co = C[:, occupied] (nbf x o),
// Cv = C[:, virtual] (nbf x v): the frozen-core / dropped-virtual slices the reference takes in Chonky.jl:38-41.
// The transformation runs on the handle's first GPU; the MO blocks then take the conventional route from device memory (a
// multi-GPU handle broadcasts them over NVLink; in rank mode every process transforms its own copy).
static int upload_ao_impl(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                          const double* Co, const double* Cv, const double* fo, const double* fv, bool sync)
{
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    const i64 n1 = nbf, n2 = n1 * nbf, n3 = n2 * nbf;
    // the copy stream continues from whatever the compute stream still has in flight
    CK(cudaEventRecord(d.ev_start, d.stream));
    CK(cudaStreamWaitEvent(d.copy, d.ev_start, 0));
    const double *dCo = Co, *dCv = Cv;
    if (classify(Co) != PK_DEVICE) {
        if (d.sCo.ensure((size_t)nbf * o * sizeof(double))) return 1;
        if (stage_now(h, d, d.sCo.p, Co, (size_t)nbf * o * sizeof(double), classify(Co))) return 1;
        dCo = d.sCo.d();
    }
    if (classify(Cv) != PK_DEVICE) {
        if (d.sCv.ensure((size_t)nbf * v * sizeof(double))) return 1;
        if (stage_now(h, d, d.sCv.p, Cv, (size_t)nbf * v * sizeof(double), classify(Cv))) return 1;
        dCv = d.sCv.d();
    }
    if (copy_then_stream(d)) return 1;
    if (d.aoQ1.ensure((size_t)n3 * o * sizeof(double))) return 1;
    // quarter 1: Q1[(nu,rho,sigma), i] = sum_mu AOERI[mu,(nu,rho,sigma)] Co[mu,i], streamed over sigma slabs of the AO tensor
    {
        const PtrKind kind = classify(AOERI);
        int schunk = nbf;
        if (kind != PK_DEVICE) {
            const size_t budget = (size_t)128 << 20;
            schunk = (int)std::max<size_t>(1, budget / ((size_t)n3 * sizeof(double)));
            if (schunk > nbf) schunk = nbf;
            for (int b = 0; b < 2; b++)
                if (d.sChunk[b].ensure((size_t)schunk * n3 * sizeof(double))) return 1;
        }
        int c = 0;
        for (int s0 = 0; s0 < nbf; s0 += schunk, c++) {
            const int sn = std::min(schunk, nbf - s0);
            const double* src = AOERI + (size_t)s0 * n3;
            if (kind != PK_DEVICE) {
                const int bsel = c & 1;
                if (c >= 2) CK(cudaStreamWaitEvent(d.copy, d.ev_free[bsel], 0));
                if (stage_now(h, d, d.sChunk[bsel].p, src, (size_t)sn * n3 * sizeof(double), kind)) return 1;
                if (copy_then_stream(d)) return 1;
                src = d.sChunk[bsel].d();
            }
            // rows (nu,rho,sigma) of this slab are rows [s0*nbf^2, (s0+sn)*nbf^2) of Q1, whose leading dimension is nbf^3
            if (quarter(h, d, d.aoQ1.d() + (size_t)s0 * n2, src, dCo, (i64)sn * n2, o, nbf, n3)) return 1;
            if (kind != PK_DEVICE) CK(cudaEventRecord(d.ev_free[c & 1], d.stream));
        }
    }
    // quarter 2: contract nu.  Q2v[(rho,sigma,i), a], Q2o[(rho,sigma,i), j]
    if (d.aoQ2v.ensure((size_t)n2 * o * v * sizeof(double))) return 1;
    if (d.aoQ2o.ensure((size_t)n2 * o * o * sizeof(double))) return 1;
    if (quarter(h, d, d.aoQ2v.d(), d.aoQ1.d(), dCv, n2 * o, v, nbf)) return 1;
    if (quarter(h, d, d.aoQ2o.d(), d.aoQ1.d(), dCo, n2 * o, o, nbf)) return 1;
    // quarter 3: contract rho.  Q3vv[(sigma,i,a), b], Q3vo[(sigma,i,a), j], Q3oo[(sigma,i,j), k]
    if (d.aoQ3vv.ensure((size_t)n1 * o * v * v * sizeof(double))) return 1;
    if (d.aoQ3vo.ensure((size_t)n1 * o * v * o * sizeof(double))) return 1;
    if (d.aoQ3oo.ensure((size_t)n1 * o * o * o * sizeof(double))) return 1;
    if (quarter(h, d, d.aoQ3vv.d(), d.aoQ2v.d(), dCv, n1 * o * v, v, nbf)) return 1;
    if (quarter(h, d, d.aoQ3vo.d(), d.aoQ2v.d(), dCo, n1 * o * v, o, nbf)) return 1;
    if (quarter(h, d, d.aoQ3oo.d(), d.aoQ2o.d(), dCo, n1 * o * o, o, nbf)) return 1;
    // quarter 4: contract sigma with Cv -> OVVV[i,a,b,c], OVOV[i,a,j,b], OOOV[i,j,k,a] in the reference's layouts
    if (d.aoOVVV.ensure((size_t)o * v * v * v * sizeof(double))) return 1;
    if (d.aoOVOV.ensure((size_t)o * v * o * v * sizeof(double))) return 1;
    if (d.aoOOOV.ensure((size_t)o * o * o * v * sizeof(double))) return 1;
    if (quarter(h, d, d.aoOVVV.d(), d.aoQ3vv.d(), dCv, (i64)o * v * v, v, nbf)) return 1;
    if (quarter(h, d, d.aoOVOV.d(), d.aoQ3vo.d(), dCv, (i64)o * v * o, v, nbf)) return 1;
    if (quarter(h, d, d.aoOOOV.d(), d.aoQ3oo.d(), dCv, (i64)o * o * o, v, nbf)) return 1;
    const int ao_launches = h->launches;
    if (upload_conv_impl(h, o, v, T1, T2, d.aoOVVV.d(), d.aoOOOV.d(), d.aoOVOV.d(), fo, fv, sync)) return 1;
    h->launches += ao_launches;
    return 0;
}

static int check_ao_args(fpt_handle* h, const char* who, int nbf, int o, int v)
{
    if (check_idle(h, who)) return 1;
    if (nbf < 1 || o < 1 || v < 1 || o + v > nbf) return fail("%s: invalid dimensions nbf=%d o=%d v=%d", who, nbf, o, v);
    return 0;
}

extern "C" int fpt_upload_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                             const double* Co, const double* Cv, const double* fo, const double* fv)
{
    if (check_ao_args(h, "fpt_upload_ao", nbf, o, v)) return 1;
    if (!T1 || !T2 || !AOERI || !Co || !Cv || !fo || !fv) return fail("fpt_upload_ao: NULL array argument");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_upload_ao", {T1, T2, AOERI, Co, Cv, fo, fv})) return 1;
    upload_begin(h);
    if (upload_ao_impl(h, nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv, true)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.upload_ms = ms_since(t0);
    return 0;
}

extern "C" int fpt_triples_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                              const double* Co, const double* Cv, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    if (check_ao_args(h, "fpt_triples_ao", nbf, o, v)) return 1;
    if (!T1 || !T2 || !AOERI || !Co || !Cv || !fo || !fv || !Et) return fail("fpt_triples_ao: NULL argument");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_triples_ao", {T1, T2, AOERI, Co, Cv, fo, fv})) return 1;
    upload_begin(h);
    if (upload_ao_impl(h, nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv, false)) return 1;
    return finish_call(h, false, t0, Et, st);
}

// Sparse AO list (the reference's default conventional container): `nint` symmetry-unique integrals, vals[z] = (mu nu|rho sigma)
// with zero-based indices idx[4z..4z+3] stored as `index_bytes`-wide integers (2: Vector{NTuple{4,Int16}}, 4: Int32).
static int upload_ao_sparse_impl(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                                 const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                                 const double* fo, const double* fv, bool sync)
{
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    const size_t n4 = (size_t)nbf * nbf * nbf * nbf;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (n4 * sizeof(double) > free_b + d.aoDense.cap)
        return fail("fpt_upload_ao_sparse: the dense AO tensor (%.1f GB for nbf=%d) does not fit the device", n4 * 8e-9, nbf);
    if (d.aoDense.ensure(n4 * sizeof(double))) return 1;
    if (d.aoFlag.ensure(sizeof(int))) return 1;
    CK(cudaMemsetAsync(d.aoDense.p, 0, n4 * sizeof(double), d.stream));
    CK(cudaMemsetAsync(d.aoFlag.p, 0, sizeof(int), d.stream));
    if (nint > 0) {
        CK(cudaEventRecord(d.ev_start, d.stream));
        CK(cudaStreamWaitEvent(d.copy, d.ev_start, 0));
        const void* didx = idx;
        const double* dvals = vals;
        if (classify(idx) != PK_DEVICE) {
            if (d.sIdx.ensure((size_t)nint * 4 * index_bytes)) return 1;
            if (stage_now(h, d, d.sIdx.p, idx, (size_t)nint * 4 * index_bytes, classify(idx))) return 1;
            didx = d.sIdx.p;
        }
        if (classify(vals) != PK_DEVICE) {
            if (d.sVals.ensure((size_t)nint * sizeof(double))) return 1;
            if (stage_now(h, d, d.sVals.p, vals, (size_t)nint * sizeof(double), classify(vals))) return 1;
            dvals = d.sVals.d();
        }
        if (copy_then_stream(d)) return 1;
        const int grid = (int)std::min<long long>((nint + 255) / 256, 148LL * 32);
        if (index_bytes == 2)
            expand_sparse_eri_kernel<short><<<grid, 256, 0, d.stream>>>(d.aoDense.d(), (const short*)didx, dvals, nint, nbf, (int*)d.aoFlag.p);
        else
            expand_sparse_eri_kernel<int><<<grid, 256, 0, d.stream>>>(d.aoDense.d(), (const int*)didx, dvals, nint, nbf, (int*)d.aoFlag.p);
        CK(cudaGetLastError());
        int bad = 0;
        CK(cudaMemcpyAsync(&bad, d.aoFlag.p, sizeof(int), cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        if (bad) return fail("fpt_upload_ao_sparse: the integral list holds an index outside [0, %d) (indices are zero-based)", nbf);
    }
    if (upload_ao_impl(h, nbf, o, v, T1, T2, d.aoDense.d(), Co, Cv, fo, fv, sync)) return 1;
    h->launches += 1;
    return 0;
}

static int check_sparse_args(fpt_handle* h, const char* who, int nbf, int o, int v, long long nint, const void* idx, int index_bytes,
                             const double* vals)
{
    if (check_ao_args(h, who, nbf, o, v)) return 1;
    if (nint < 0 || (nint > 0 && (!idx || !vals))) return fail("%s: invalid integral list (nint=%lld)", who, nint);
    if (index_bytes != 2 && index_bytes != 4) return fail("%s: index_bytes must be 2 or 4, got %d", who, index_bytes);
    if (index_bytes == 2 && nbf > 32767) return fail("%s: nbf=%d does not fit 16-bit indices", who, nbf);
    return 0;
}

extern "C" int fpt_upload_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                                    const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                                    const double* fo, const double* fv)
{
    if (check_sparse_args(h, "fpt_upload_ao_sparse", nbf, o, v, nint, idx, index_bytes, vals)) return 1;
    if (!T1 || !T2 || !Co || !Cv || !fo || !fv) return fail("fpt_upload_ao_sparse: NULL array argument");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_upload_ao_sparse", {T1, T2, idx, vals, Co, Cv, fo, fv})) return 1;
    upload_begin(h);
    if (upload_ao_sparse_impl(h, nbf, o, v, T1, T2, nint, idx, index_bytes, vals, Co, Cv, fo, fv, true)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    h->last.upload_ms = ms_since(t0);
    return 0;
}

extern "C" int fpt_triples_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                                     const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                                     const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    if (check_sparse_args(h, "fpt_triples_ao_sparse", nbf, o, v, nint, idx, index_bytes, vals)) return 1;
    if (!T1 || !T2 || !Co || !Cv || !fo || !fv || !Et) return fail("fpt_triples_ao_sparse: NULL argument");
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, "fpt_triples_ao_sparse", {T1, T2, idx, vals, Co, Cv, fo, fv})) return 1;
    upload_begin(h);
    if (upload_ao_sparse_impl(h, nbf, o, v, T1, T2, nint, idx, index_bytes, vals, Co, Cv, fo, fv, false)) return 1;
    return finish_call(h, false, t0, Et, st);
}

// ---- calibration and diagnostics ---------------------------------------------------------------------------------------------
extern "C" int fpt_fp64_peak(fpt_handle* h, int variant, double ms_target, double* tflops)
{
    if (!h || !tflops) return fail("fpt_fp64_peak: NULL argument");
    if (check_idle(h, "fpt_fp64_peak")) return 1;
    DeviceGuard guard;
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    if (d.out.ensure(sizeof(double))) return 1;
    const int iters = 4096;
    const int grid = d.n_sm * 8;   // 8 CTAs x 8 warps per SM -> 16 warps per SMSP
    // flops per launch
    const double fl = (variant == 0) ? (double)grid * 8 /*warps*/ * iters * 16.0 * 512.0
                                     : (double)grid * 256 /*threads*/ * iters * 16.0 * 2.0;
    auto launch = [&]() {
        if (variant == 0) peak_dmma_kernel<<<grid, 256, 0, d.stream>>>(d.out.d(), iters, 1e-3);
        else peak_dfma_kernel<<<grid, 256, 0, d.stream>>>(d.out.d(), iters, 1e-3);
    };
    launch();
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaGetLastError());
    // calibrate launch count
    CK(cudaEventRecord(d.ev0[0], d.stream));
    launch();
    CK(cudaEventRecord(d.ev1[0], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    float ms1 = 0.f;
    CK(cudaEventElapsedTime(&ms1, d.ev0[0], d.ev1[0]));
    int reps = (int)(ms_target / (ms1 > 1e-3f ? ms1 : 1e-3f));
    if (reps < 1) reps = 1;
    if (reps > 20000) reps = 20000;
    CK(cudaEventRecord(d.ev0[0], d.stream));
    for (int t = 0; t < reps; t++) launch();
    CK(cudaEventRecord(d.ev1[0], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, d.ev0[0], d.ev1[0]));
    *tflops = fl * reps / (ms * 1e-3) / 1e12;
    return 0;
}

// Stand-alone timing of the K3 / K5 GEMM on synthetic operands (measurement aid): C(M x N) = A(M x K) . B(N x K)^T, column-major
// output, `reps` launches; returns the sustained TFLOP/s (2 M N K per launch).
extern "C" int fpt_gemm_bench(fpt_handle* h, long long M, int N, int K, int reps, double* tflops)
{
    if (!h || !tflops) return fail("fpt_gemm_bench: NULL argument");
    if (check_idle(h, "fpt_gemm_bench")) return 1;
    if (M < 1 || N < 1 || K < 1 || reps < 1) return fail("fpt_gemm_bench: invalid shape");
    DeviceGuard guard;
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    DevBuf A, B, C;
    if (A.ensure((size_t)M * K * sizeof(double)) || B.ensure((size_t)N * K * sizeof(double)) || C.ensure((size_t)M * N * sizeof(double))) {
        A.release(); B.release(); C.release();
        return 1;
    }
    cudaMemsetAsync(A.p, 0, (size_t)M * K * sizeof(double), d.stream);
    cudaMemsetAsync(B.p, 0, (size_t)N * K * sizeof(double), d.stream);
    GemmOut out{};
    out.C = C.d();
    out.ldc = M;
    cudaError_t e = gemm_tn_launch<EPI_COLMAJOR>(d.stream, A.d(), rowmap_identity(), B.d(), rowmap_identity(), M, N, K, out);
    cudaEventRecord(d.ev0[0], d.stream);
    for (int r = 0; r < reps && e == cudaSuccess; r++)
        e = gemm_tn_launch<EPI_COLMAJOR>(d.stream, A.d(), rowmap_identity(), B.d(), rowmap_identity(), M, N, K, out);
    cudaEventRecord(d.ev1[0], d.stream);
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, d.ev0[0], d.ev1[0]);
    A.release(); B.release(); C.release();
    if (e != cudaSuccess || e2 != cudaSuccess) return fail("fpt_gemm_bench: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    *tflops = 2.0 * (double)M * N * K * reps / (ms * 1e-3) / 1e12;
    return 0;
}

extern "C" int fpt_set_debug_flags(fpt_handle* h, int flags)
{
    if (!h) return fail("fpt_set_debug_flags: NULL handle");
    h->dbg_flags = flags;   // read by every GPU of the handle at the next compute
    return 0;
}

// Restrict the work list to positions [t_begin, t_end) of the reference's flattened i >= j >= k triplet list (k fastest,
// zero-weight i = j = k entries included, exactly the list the loops of ijk.jl:49,63,83 walk); t_end < 0 = to the end.
extern "C" int fpt_set_triplet_window(fpt_handle* h, long long t_begin, long long t_end)
{
    if (!h) return fail("fpt_set_triplet_window: NULL handle");
    if (!h->loaded) return fail("fpt_set_triplet_window: no problem uploaded");
    const i64 nfull = (i64)h->o * (h->o + 1) * (h->o + 2) / 6;
    if (t_end < 0 || t_end > nfull) t_end = nfull;
    if (t_begin < 0) t_begin = 0;
    if (t_begin > t_end) t_begin = t_end;
    const i64 u0 = triplets_before(h->o, t_begin), u1 = triplets_before(h->o, t_end);
    h->tw_begin = u0;
    h->tw_count = u1 - u0;
    h->nitems = h->devs[0]->prob.nb * h->tw_count;
    return 0;
}

// 1: block-major (default), 0: triplet-major.  Takes effect for the next compute; keeps the triplet window.
extern "C" int fpt_set_item_order(fpt_handle* h, int order)
{
    if (!h) return fail("fpt_set_item_order: NULL handle");
    if (order != 0 && order != 1) return fail("fpt_set_item_order: order must be 0 or 1, got %d", order);
    h->item_order = order;
    return 0;
}

// Part `rank` of `world` of the current work list, as an item range for fpt_compute: contiguous, equal estimated cost.
extern "C" int fpt_shard_items(fpt_handle* h, int rank, int world, long long* item_begin, long long* item_end)
{
    if (!h || !item_begin || !item_end) return fail("fpt_shard_items: NULL argument");
    if (!h->loaded) return fail("fpt_shard_items: no problem uploaded");
    if (world < 1 || rank < 0 || rank >= world) return fail("fpt_shard_items: invalid rank %d of %d", rank, world);
    i64 sb, se;
    shard_range(h, current_problem(h, *h->devs[0]), 0, h->nitems, rank, world, &sb, &se);
    *item_begin = sb;
    *item_end = se;
    return 0;
}

extern "C" int fpt_set_kernel_variant(fpt_handle* h, int variant)
{
    if (!h) return fail("fpt_set_kernel_variant: NULL handle");
#ifdef FPT_WITH_VARIANT2
    if (variant != 1 && variant != 2) return fail("fpt_set_kernel_variant: variant must be 1 or 2, got %d", variant);
#else
    if (variant != 1) return fail("fpt_set_kernel_variant: variant %d is not in this build (the experimental epilogue-warp kernel needs -DFPT_WITH_VARIANT2)", variant);
#endif
    h->kernel_variant = variant;
    return 0;
}

extern "C" int fpt_set_profiling(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_profiling: NULL handle");
    h->profiling = on != 0;
    return 0;
}

// Phase breakdown of the last fpt_compute on the handle's first GPU (cycles summed over CTAs; see the header for the 24 entries)
extern "C" int fpt_last_profile(fpt_handle* h, double* out24)
{
    if (!h || !out24) return fail("fpt_last_profile: NULL argument");
    Dev& d = *h->devs[0];
    if (d.last_grid <= 0 || !h->last_profiled) return fail("fpt_last_profile: the last compute was not profiled (fpt_set_profiling)");
    DeviceGuard guard;
    CK(cudaSetDevice(d.dev));
    std::vector<long long> buf((size_t)d.last_grid * NPROF);
    CK(cudaMemcpy(buf.data(), d.prof.p, buf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int t = 0; t < NPROF; t++) out24[t] = 0.0;
    for (int b = 0; b < d.last_grid; b++)
        for (int t = 0; t < NPROF; t++) out24[t] += (double)buf[(size_t)b * NPROF + t];
    return 0;
}

// DMMA issue study (design aid): sustained TFLOP/s with `ilp` independent accumulators per warp and
// `warps_per_sm` warps on each SM (1 CTA/SM).
extern "C" int fpt_dmma_sweep(fpt_handle* h, int ilp, int warps_per_sm, double* tflops)
{
    if (!h || !tflops) return fail("fpt_dmma_sweep: NULL argument");
    if (check_idle(h, "fpt_dmma_sweep")) return 1;
    DeviceGuard guard;
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    if (d.out.ensure(sizeof(double))) return 1;
    const int iters = 20000 / ilp;
    const int threads = warps_per_sm * 32;
    if (threads < 32 || threads > 1024) return fail("fpt_dmma_sweep: warps_per_sm out of range");
    void (*k)(double*, int, double) = nullptr;
    switch (ilp) {
    case 1: k = dmma_ilp_kernel<1>; break;
    case 2: k = dmma_ilp_kernel<2>; break;
    case 4: k = dmma_ilp_kernel<4>; break;
    case 8: k = dmma_ilp_kernel<8>; break;
    case 16: k = dmma_ilp_kernel<16>; break;
    case 32: k = dmma_ilp_kernel<32>; break;
    default: return fail("fpt_dmma_sweep: ilp must be 1,2,4,8,16,32");
    }
    const size_t smem = 120 * 1024;   // > half of the SM: forces 1 CTA/SM
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<d.n_sm, threads, smem, d.stream>>>(d.out.d(), iters, 1e-3);
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaEventRecord(d.ev0[0], d.stream));
    for (int r = 0; r < 5; r++) k<<<d.n_sm, threads, smem, d.stream>>>(d.out.d(), iters, 1e-3);
    CK(cudaEventRecord(d.ev1[0], d.stream));
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, d.ev0[0], d.ev1[0]));
    *tflops = 5.0 * d.n_sm * warps_per_sm * (double)iters * ilp * 512.0 / (ms * 1e-3) / 1e12;
    return 0;
}
