// C-ABI of libfermi_pt_b200.so (declared in include/fermi_pt_b200.h) and the host-side driver of the (T) path:
// device buffer pool, host->device staging, layout prep / DF assembly launches, the persistent fused kernel launch
// and the final reduction.  Replaces the driver loop + accumulation of
//   src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150 (reference, Julia threads)
// No CPU fallback: every entry point fails loudly without a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fermi_pt_b200.h"
#include "fpt_aux_kernels.cuh"
#include "fpt_triples.cuh"
#include "fpt_triples2.cuh"

using namespace fpt;

static thread_local std::string g_err;
static int fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    double* d() const { return (double*)p; }
};

struct fpt_handle {
    int dev = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // resident operands
    DevBuf Pt, Qt, OV2, T1d, fo, fv, prefix, partials, counter, out, prof, blocktab;
    // staging for raw inputs
    DevBuf sT1, sT2, sOOOV, sOVOV, sChunk, sBOO, sBOV, sBVV;
    // AO -> MO route: coefficient blocks, quarter-transformed intermediates, the MO blocks the (T) path consumes
    DevBuf sCo, sCv, aoDense, sIdx, sVals, aoQ1, aoQ2v, aoQ2o, aoQ3vv, aoQ3vo, aoQ3oo, aoOVVV, aoOOOV, aoOVOV;
    Problem prob{};
    bool loaded = false;
    fpt_stats last{};
    int launches = 0;
    int last_grid = 0;
    bool profiling = false;
    int dbg_flags = 0;
    int item_order = 1;       // 1: block-major (default), 0: triplet-major (see Problem::order)
    std::vector<double> block_cost;
    int kernel_variant = 1;   // 1: DMMA warps add their own accumulators into the W slots (fpt_triples.cuh, default);
                              // 2: experimental epilogue-warp kernel with TMEM parking (fpt_triples2.cuh)
    bool last_profiled = false;
    // multi-GPU (single process): this handle drives devices[0]; peers[] drive the others; one NCCL clique
    std::vector<fpt_handle*> peers;
    std::vector<ncclComm_t> comms;   // comms[0] = this device, comms[1+k] = peers[k]
};

// ---- NCCL, resolved at run time (only handles with ngpu > 1 need it) ---------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail("multi-GPU handle: cannot load libnccl.so.2 (%s)", dlerror());
#define FPT_SYM(field, name)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));          \
    if (!g_nccl.field) return fail("multi-GPU handle: libnccl lacks %s", name)
    FPT_SYM(CommInitAll, "ncclCommInitAll");
    FPT_SYM(CommDestroy, "ncclCommDestroy");
    FPT_SYM(GroupStart, "ncclGroupStart");
    FPT_SYM(GroupEnd, "ncclGroupEnd");
    FPT_SYM(Broadcast, "ncclBroadcast");
    FPT_SYM(AllReduce, "ncclAllReduce");
    FPT_SYM(GetErrorString, "ncclGetErrorString");
#undef FPT_SYM
    g_nccl.lib = lib;
    return 0;
}
#define NCK(call)                                                                                              \
    do {                                                                                                       \
        ncclResult_t r_ = (call);                                                                              \
        if (r_ != ncclSuccess) return fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString(r_)); \
    } while (0)

static int broadcast_operands(fpt_handle* h);

extern "C" const char* fpt_last_error(void) { return g_err.c_str(); }
extern "C" const char* fpt_version(void) { return "fermi_pt_b200 0.3 (sm_100a)"; }

static int create_one(int dev, fpt_handle** out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("fpt_create: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail("fpt_create: device %d out of range (have %d)", dev, ndev);
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail("fpt_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    fpt_handle* h = new fpt_handle();
    h->dev = dev;
    h->n_sm = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaFuncSetAttribute(triples_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
    *out = h;
    return 0;
}

extern "C" int fpt_create(int ngpu, const int* devices, fpt_handle** out)
{
    if (!out) return fail("fpt_create: out is NULL");
    *out = nullptr;
    if (ngpu < 1 || ngpu > 16) return fail("fpt_create: ngpu=%d out of range", ngpu);
    if (ngpu > 1 && !devices) return fail("fpt_create: a device list is required for ngpu > 1");
    fpt_handle* h = nullptr;
    if (create_one(devices ? devices[0] : -1, &h)) return 1;
    if (ngpu > 1) {
        // single-process multi-GPU: one NCCL clique over NVLink, operands broadcast once per upload, one scalar all-reduce
        if (nccl_load()) { fpt_destroy(h); return 1; }
        for (int k = 1; k < ngpu; k++) {
            fpt_handle* p = nullptr;
            if (create_one(devices[k], &p)) { fpt_destroy(h); return 1; }
            h->peers.push_back(p);
        }
        h->comms.resize(ngpu);
        ncclResult_t r = g_nccl.CommInitAll(h->comms.data(), ngpu, devices);
        if (r != ncclSuccess) {
            h->comms.clear();
            fail("ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
            fpt_destroy(h);
            return 1;
        }
        CK(cudaSetDevice(h->dev));
    }
    *out = h;
    return 0;
}

extern "C" int fpt_destroy(fpt_handle* h)
{
    if (!h) return 0;
    for (ncclComm_t c : h->comms) g_nccl.CommDestroy(c);
    h->comms.clear();
    for (fpt_handle* p : h->peers) fpt_destroy(p);
    h->peers.clear();
    cudaSetDevice(h->dev);
    DevBuf* bufs[] = {&h->Pt, &h->Qt, &h->OV2, &h->T1d, &h->fo, &h->fv, &h->prefix, &h->partials, &h->counter, &h->out, &h->prof, &h->blocktab,
                      &h->sCo, &h->sCv, &h->aoDense, &h->sIdx, &h->sVals, &h->aoQ1, &h->aoQ2v, &h->aoQ2o, &h->aoQ3vv, &h->aoQ3vo, &h->aoQ3oo, &h->aoOVVV, &h->aoOOOV, &h->aoOVOV, &h->sT1, &h->sT2, &h->sOOOV, &h->sOVOV, &h->sChunk, &h->sBOO, &h->sBOV, &h->sBVV};
    for (DevBuf* b : bufs) b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

static bool is_device_ptr(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Returns a device pointer for `src` (n doubles): src itself if already on the device, else a staged copy.
static int stage_in(fpt_handle* h, DevBuf& buf, const double* src, size_t n, const double** dptr, double* h2d_bytes)
{
    if (is_device_ptr(src)) { *dptr = src; return 0; }
    if (buf.ensure(n * sizeof(double))) return 1;
    CK(cudaMemcpyAsync(buf.p, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    *h2d_bytes += (double)n * sizeof(double);
    *dptr = buf.d();
    return 0;
}

static int setup_problem(fpt_handle* h, int o, int v)
{
    if (o < 1 || v < 1) return fail("invalid dimensions o=%d v=%d", o, v);
    Problem& P = h->prob;
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
    if (h->Pt.ensure((size_t)o * P.vp * P.vp * P.Kp * sizeof(double))) return 1;
    if (h->Qt.ensure((size_t)o * o * P.G * P.vp * KGROUP * sizeof(double))) return 1;
    if (h->OV2.ensure((size_t)ov2_elems(P) * sizeof(double))) return 1;
    if (h->T1d.ensure((size_t)o * v * sizeof(double))) return 1;
    if (h->fo.ensure((size_t)o * sizeof(double))) return 1;
    if (h->fv.ensure((size_t)v * sizeof(double))) return 1;
    if (h->partials.ensure((size_t)h->n_sm * 4 * sizeof(double))) return 1;
    if (h->counter.ensure(sizeof(unsigned long long))) return 1;
    if (h->out.ensure(sizeof(double))) return 1;
    if (h->prof.ensure((size_t)h->n_sm * NPROF * sizeof(long long))) return 1;
    // block descriptor table (positions in (i,j,k) instead of orbital numbers)
    std::vector<BlockTabEntry> tab((size_t)P.nb);
    for (i64 b = 0; b < P.nb; b++) {
        int A, B, C;
        tetra_decode(b, A, B, C);
        make_block(A, B, C, P.vp, tab[b].bd);
        tab[b].ngemm = make_gemms(tab[b].bd, 0, 1, 2, tab[b].gemm);
        make_fast_order(tab[b]);
    }
    if (h->blocktab.ensure(tab.size() * sizeof(BlockTabEntry))) return 1;
    CK(cudaMemcpyAsync(h->blocktab.p, tab.data(), tab.size() * sizeof(BlockTabEntry), cudaMemcpyHostToDevice, h->stream));
    P.blocktab = (const BlockTabEntry*)h->blocktab.p;
    CK(cudaStreamSynchronize(h->stream));   // `tab` is a local
    h->block_cost.resize((size_t)P.nb);
    for (i64 b = 0; b < P.nb; b++) h->block_cost[b] = block_cost(tab[b], P.G);
    P.Pt = h->Pt.d(); P.Qt = h->Qt.d(); P.OV2 = h->OV2.d(); P.T1d = h->T1d.d();
    P.fo = h->fo.d(); P.fv = h->fv.d();
    return 0;
}

static int grid1d(i64 n, int block = 256) { i64 g = (n + block - 1) / block; if (g > 148 * 32) g = 148 * 32; if (g < 1) g = 1; return (int)g; }

// the parts common to conventional and DF uploads: T1, T2 -> T1d, Pt hole part, Qt particle part; fo, fv
static int upload_common(fpt_handle* h, const double* T1, const double* T2, const double* fo, const double* fv,
                         const double** dT2, double* h2d)
{
    const Problem& P = h->prob;
    const int o = P.o, v = P.v;
    const double* dT1;
    if (stage_in(h, h->sT1, T1, (size_t)o * v, &dT1, h2d)) return 1;
    if (stage_in(h, h->sT2, T2, (size_t)o * o * v * v, dT2, h2d)) return 1;
    CK(cudaMemcpyAsync(h->fo.p, fo, o * sizeof(double), cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(h->fv.p, fv, v * sizeof(double), cudaMemcpyDefault, h->stream));
    if (!is_device_ptr(fo)) *h2d += (o + v) * sizeof(double);
    CK(cudaMemsetAsync(h->Pt.p, 0, (size_t)o * P.vp * P.vp * P.Kp * sizeof(double), h->stream));
    prep_t1<<<grid1d(o * v), 256, 0, h->stream>>>(P, h->T1d.d(), dT1);
    prep_pt_hole<<<grid1d((i64)o * o * v * v), 256, 0, h->stream>>>(P, h->Pt.d(), *dT2);
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int fpt_upload_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                               const double* OOOV, const double* OVOV, const double* fo, const double* fv)
{
    if (!h) return fail("fpt_upload_conv: NULL handle");
    if (!T1 || !T2 || !OVVV || !OOOV || !OVOV || !fo || !fv) return fail("fpt_upload_conv: NULL array argument");
    CK(cudaSetDevice(h->dev));
    auto t0 = std::chrono::steady_clock::now();
    h->loaded = false;
    h->launches = 0;
    if (setup_problem(h, o, v)) return 1;
    const Problem& P = h->prob;
    double h2d = 0.0;
    const double* dT2;
    if (upload_common(h, T1, T2, fo, fv, &dT2, &h2d)) return 1;
    const double *dOOOV, *dOVOV;
    if (stage_in(h, h->sOOOV, OOOV, (size_t)o * o * o * v, &dOOOV, &h2d)) return 1;
    if (stage_in(h, h->sOVOV, OVOV, (size_t)o * v * o * v, &dOVOV, &h2d)) return 1;
    prep_qt<<<grid1d((i64)o * o * P.G * P.vp * KGROUP), 256, 0, h->stream>>>(P, h->Qt.d(), dT2, dOOOV);
    prep_ov2<<<grid1d(ov2_elems(P)), 256, 0, h->stream>>>(P, h->OV2.d(), dOVOV);
    h->launches += 2;
    CK(cudaGetLastError());
    // OVVV -> Pt particle part, in chunks over the slowest index d
    const size_t slab = (size_t)o * v * v;   // doubles per d
    const bool on_dev = is_device_ptr(OVVV);
    int dchunk = v;
    if (!on_dev) {
        const size_t budget = (size_t)512 << 20;
        dchunk = (int)(budget / (slab * sizeof(double)));
        if (dchunk < 1) dchunk = 1;
        if (dchunk > v) dchunk = v;
        if (h->sChunk.ensure((size_t)dchunk * slab * sizeof(double))) return 1;
    }
    for (int d0 = 0; d0 < v; d0 += dchunk) {
        const int dn = (v - d0 < dchunk) ? v - d0 : dchunk;
        const double* src = OVVV + (size_t)d0 * slab;
        if (!on_dev) {
            CK(cudaMemcpyAsync(h->sChunk.p, src, (size_t)dn * slab * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            h2d += (double)dn * slab * sizeof(double);
            src = h->sChunk.d();
        }
        dim3 grid((unsigned)(((size_t)o * v + 31) / 32), (unsigned)((dn + 31) / 32), (unsigned)v);
        prep_pt_particle<<<grid, dim3(32, 8), 0, h->stream>>>(P, h->Pt.d(), src, d0, dn);
        h->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    h->loaded = true;
    if (!h->peers.empty() && broadcast_operands(h)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h2d;
    h->last.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

extern "C" int fpt_upload_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                             const double* BOV, const double* BVV, const double* fo, const double* fv)
{
    if (!h) return fail("fpt_upload_df: NULL handle");
    if (!T1 || !T2 || !BOO || !BOV || !BVV || !fo || !fv) return fail("fpt_upload_df: NULL array argument");
    if (naux < 1) return fail("fpt_upload_df: invalid naux=%d", naux);
    CK(cudaSetDevice(h->dev));
    auto t0 = std::chrono::steady_clock::now();
    h->loaded = false;
    h->launches = 0;
    if (setup_problem(h, o, v)) return 1;
    const Problem& P = h->prob;
    double h2d = 0.0;
    const double* dT2;
    if (upload_common(h, T1, T2, fo, fv, &dT2, &h2d)) return 1;
    const double *dBOO, *dBOV, *dBVV;
    if (stage_in(h, h->sBOO, BOO, (size_t)naux * o * o, &dBOO, &h2d)) return 1;
    if (stage_in(h, h->sBOV, BOV, (size_t)naux * o * v, &dBOV, &h2d)) return 1;
    if (stage_in(h, h->sBVV, BVV, (size_t)naux * v * v, &dBVV, &h2d)) return 1;
    // Qt: T2 part by the gather kernel (OOOV argument unused for kappa >= v when we overwrite below) -> pass a
    // zero-filled dummy?  Instead: prep_qt with OOOV = nullptr is not allowed, so build the T2 part with a DF-aware call:
    // stage 1 fills everything (hole part from a temporary zero source is avoided by the kernel's branch order).
    if (h->sOOOV.ensure((size_t)o * o * o * v * sizeof(double))) return 1;
    CK(cudaMemsetAsync(h->sOOOV.p, 0, (size_t)o * o * o * v * sizeof(double), h->stream));
    prep_qt<<<grid1d((i64)o * o * P.G * P.vp * KGROUP), 256, 0, h->stream>>>(P, h->Qt.d(), dT2, h->sOOOV.d());
    {   // OOOV[l,q,r,z] = sum_Q BOO[Q,l,q] BOV[Q,r,z]  -> Qt hole part        (DFERI.jl:88-112)
        const int M = o * o, N = o * v;
        dim3 grid((M + 63) / 64, (N + 63) / 64);
        df_gemm_kernel<1><<<grid, 128, 0, h->stream>>>(P, h->Qt.d(), dBOO, dBOV, M, N, naux);
    }
    CK(cudaMemsetAsync(h->OV2.p, 0, (size_t)ov2_elems(P) * sizeof(double), h->stream));
    {   // OVOV[q,y,r,z] = sum_Q BOV[Q,q,y] BOV[Q,r,z]  -> OV2                 (DFERI.jl:139-154)
        const int M = o * v, N = o * v;
        dim3 grid((M + 63) / 64, (N + 63) / 64);
        df_gemm_kernel<2><<<grid, 128, 0, h->stream>>>(P, h->OV2.d(), dBOV, dBOV, M, N, naux);
    }
    {   // OVVV[p,y,x,d] = sum_Q BOV[Q,p,y] BVV[Q,x,d]  -> Pt particle part    (DFERI.jl:156-180, never on the host)
        const int M = o * v, N = v * v;
        dim3 grid((M + 63) / 64, (N + 63) / 64);
        df_gemm_kernel<0><<<grid, 128, 0, h->stream>>>(P, h->Pt.d(), dBOV, dBVV, M, N, naux);
    }
    h->launches += 4;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->loaded = true;
    if (!h->peers.empty() && broadcast_operands(h)) return 1;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h2d;
    h->last.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

extern "C" int fpt_num_items(fpt_handle* h, long long* n)
{
    if (!h || !n) return fail("fpt_num_items: NULL argument");
    if (!h->loaded) return fail("fpt_num_items: no problem uploaded");
    *n = h->prob.nitems;
    return 0;
}

// Static split of the item range [b, e) into `world` contiguous parts of equal estimated cost (shard_items in fpt_layout.h,
// shared with the CPU emulator so that the gloo tests exercise the very same split)
static void shard_range(const fpt_handle* h, i64 b, i64 e, int rank, int world, i64* sb, i64* se)
{
    shard_items(h->prob, h->block_cost.data(), b, e, rank, world, sb, se);
}

// launch the fused kernel + reduction for [item_begin, item_end) on h's device (asynchronous; result in h->out)
static int compute_launch(fpt_handle* h, i64 item_begin, i64 item_end)
{
    CK(cudaSetDevice(h->dev));
    const Problem& P = h->prob;
    const i64 n = item_end - item_begin;
    int grid = h->n_sm;
    if ((i64)grid > n) grid = (int)(n > 0 ? n : 1);
    CK(cudaMemsetAsync(h->counter.p, 0, sizeof(unsigned long long), h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    if (h->kernel_variant == 2) {
        if (h->profiling)
            triples_kernel2<true><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, h->stream>>>(
                P, item_begin, item_end, (unsigned long long*)h->counter.p, h->partials.d(), (long long*)h->prof.p);
        else
            triples_kernel2<false><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, h->stream>>>(
                P, item_begin, item_end, (unsigned long long*)h->counter.p, h->partials.d(), (long long*)h->prof.p);
    } else if (h->profiling)
        triples_kernel<true><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, h->stream>>>(
            P, item_begin, item_end, (unsigned long long*)h->counter.p, h->partials.d(), (long long*)h->prof.p);
    else
        triples_kernel<false><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, h->stream>>>(
            P, item_begin, item_end, (unsigned long long*)h->counter.p, h->partials.d(), (long long*)h->prof.p);
    h->last_grid = grid;
    h->last_profiled = h->profiling;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    reduce_partials<<<1, 32, 0, h->stream>>>(h->partials.d(), grid, h->out.d());
    CK(cudaGetLastError());
    return 0;
}

extern "C" int fpt_compute(fpt_handle* h, long long item_begin, long long item_end, double* Et, fpt_stats* st)
{
    if (!h || !Et) return fail("fpt_compute: NULL argument");
    if (!h->loaded) return fail("fpt_compute: no problem uploaded");
    const Problem& P = h->prob;
    if (item_end < 0 || item_end > P.nitems) item_end = P.nitems;
    if (item_begin < 0) item_begin = 0;
    if (item_begin > item_end) item_begin = item_end;
    const i64 n = item_end - item_begin;
    const int ng = 1 + (int)h->peers.size();
    std::vector<fpt_handle*> hs(1, h);
    for (fpt_handle* p : h->peers) hs.push_back(p);
    // static contiguous shards of the range, one per GPU (equal estimated cost)
    for (int d = 0; d < ng; d++) {
        hs[d]->profiling = h->profiling;
        hs[d]->kernel_variant = h->kernel_variant;
        hs[d]->prob.order = P.order; hs[d]->prob.tw_begin = P.tw_begin; hs[d]->prob.tw_count = P.tw_count; hs[d]->prob.nitems = P.nitems;
        i64 sb, se;
        shard_range(h, item_begin, item_end, d, ng, &sb, &se);
        if (compute_launch(hs[d], sb, se)) return 1;
    }
    if (ng > 1) {   // the single scalar all-reduce of E(T)
        NCK(g_nccl.GroupStart());
        for (int d = 0; d < ng; d++)
            NCK(g_nccl.AllReduce(hs[d]->out.p, hs[d]->out.p, 1, ncclDouble, ncclSum, h->comms[d], hs[d]->stream));
        NCK(g_nccl.GroupEnd());
    }
    double e = 0.0;
    float ms_max = 0.f;
    CK(cudaSetDevice(h->dev));
    CK(cudaMemcpyAsync(&e, h->out.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    for (int d = 0; d < ng; d++) {
        CK(cudaSetDevice(hs[d]->dev));
        CK(cudaStreamSynchronize(hs[d]->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, hs[d]->ev0, hs[d]->ev1));
        if (ms > ms_max) ms_max = ms;
    }
    CK(cudaSetDevice(h->dev));
    *Et = e;
    // algorithmic flops of the triplets in the window, scaled by the share of the window's items that were computed
    const double ntrip = (double)P.tw_count;
    h->last.kernel_ms = ms_max;
    h->last.n_items = n;
    h->last.n_triplets = (long long)ntrip;
    h->last.flops = 12.0 * P.v * (double)P.v * P.v * (P.v + P.o) * ntrip * (P.nitems ? (double)n / (double)P.nitems : 0.0);
    h->last.n_launches = h->launches + 2 * ng;
    h->last.n_sm = h->n_sm;
    if (st) *st = h->last;
    return 0;
}

// multi-GPU handle: make the prepared operands resident on every peer (one NCCL broadcast per buffer, root = device 0)
static int broadcast_operands(fpt_handle* h)
{
    const Problem& P = h->prob;
    const int ng = 1 + (int)h->peers.size();
    for (fpt_handle* p : h->peers) {
        CK(cudaSetDevice(p->dev));
        p->dbg_flags = h->dbg_flags;
        p->item_order = h->item_order;
        p->loaded = false;
        if (setup_problem(p, P.o, P.v)) return 1;
    }
    const size_t counts[6] = {(size_t)P.o * P.vp * P.vp * P.Kp, (size_t)P.o * P.o * P.G * P.vp * KGROUP, (size_t)ov2_elems(P),
                              (size_t)P.o * P.v, (size_t)P.o, (size_t)P.v};
    NCK(g_nccl.GroupStart());
    for (int d = 0; d < ng; d++) {
        fpt_handle* hd = d ? h->peers[d - 1] : h;
        void* bufs[6] = {hd->Pt.p, hd->Qt.p, hd->OV2.p, hd->T1d.p, hd->fo.p, hd->fv.p};
        for (int b = 0; b < 6; b++)
            NCK(g_nccl.Broadcast(bufs[b], bufs[b], counts[b], ncclDouble, 0, h->comms[d], hd->stream));
    }
    NCK(g_nccl.GroupEnd());
    for (int d = 0; d < ng; d++) {
        fpt_handle* hd = d ? h->peers[d - 1] : h;
        CK(cudaSetDevice(hd->dev));
        CK(cudaStreamSynchronize(hd->stream));
        hd->loaded = true;
    }
    CK(cudaSetDevice(h->dev));
    return 0;
}

extern "C" int fpt_triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                                const double* OOOV, const double* OVOV, const double* fo, const double* fv, double* Et,
                                fpt_stats* st)
{
    auto t0 = std::chrono::steady_clock::now();
    if (fpt_upload_conv(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv)) return 1;
    if (fpt_compute(h, 0, -1, Et, nullptr)) return 1;
    h->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (st) *st = h->last;
    return 0;
}

extern "C" int fpt_triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                              const double* BOV, const double* BVV, const double* fo, const double* fv, double* Et,
                              fpt_stats* st)
{
    auto t0 = std::chrono::steady_clock::now();
    if (fpt_upload_df(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv)) return 1;
    if (fpt_compute(h, 0, -1, Et, nullptr)) return 1;
    h->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (st) *st = h->last;
    return 0;
}

// ---- AO -> MO route (replaces Chonky.jl:28-114 for the three blocks the (T) path reads) --------------------------
static int quarter(fpt_handle* h, double* C, const double* A, const double* B, i64 M, int N, int Q, i64 ldc = 0)
{
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    quarter_gemm_kernel<<<grid, 128, 0, h->stream>>>(C, A, B, M, N, Q, ldc ? ldc : M);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// AOERI[mu,nu,rho,sigma] (nbf^4, column-major, chemist notation as in aoints["ERI"]), Co = C[:, occupied] (nbf x o),
// Cv = C[:, virtual] (nbf x v): the frozen-core / dropped-virtual slices the reference takes in Chonky.jl:38-41.
extern "C" int fpt_upload_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                             const double* Co, const double* Cv, const double* fo, const double* fv)
{
    if (!h) return fail("fpt_upload_ao: NULL handle");
    if (!T1 || !T2 || !AOERI || !Co || !Cv || !fo || !fv) return fail("fpt_upload_ao: NULL array argument");
    if (nbf < 1 || o < 1 || v < 1 || o + v > nbf) return fail("fpt_upload_ao: invalid dimensions nbf=%d o=%d v=%d", nbf, o, v);
    CK(cudaSetDevice(h->dev));
    auto t0 = std::chrono::steady_clock::now();
    h->loaded = false;
    h->launches = 0;
    double h2d = 0.0;
    const i64 n1 = nbf, n2 = n1 * nbf, n3 = n2 * nbf;
    const double *dCo, *dCv;
    if (stage_in(h, h->sCo, Co, (size_t)nbf * o, &dCo, &h2d)) return 1;
    if (stage_in(h, h->sCv, Cv, (size_t)nbf * v, &dCv, &h2d)) return 1;
    if (h->aoQ1.ensure((size_t)n3 * o * sizeof(double))) return 1;
    // quarter 1: Q1[(nu,rho,sigma), i] = sum_mu AOERI[mu,(nu,rho,sigma)] Co[mu,i], streamed over sigma slabs of the AO tensor
    {
        const bool on_dev = is_device_ptr(AOERI);
        int schunk = nbf;
        if (!on_dev) {
            const size_t budget = (size_t)512 << 20;
            schunk = (int)(budget / ((size_t)n3 * sizeof(double)));
            if (schunk < 1) schunk = 1;
            if (schunk > nbf) schunk = nbf;
            if (h->sChunk.ensure((size_t)schunk * n3 * sizeof(double))) return 1;
        }
        for (int s0 = 0; s0 < nbf; s0 += schunk) {
            const int sn = (nbf - s0 < schunk) ? nbf - s0 : schunk;
            const double* src = AOERI + (size_t)s0 * n3;
            if (!on_dev) {
                CK(cudaMemcpyAsync(h->sChunk.p, src, (size_t)sn * n3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                h2d += (double)sn * n3 * sizeof(double);
                src = h->sChunk.d();
            }
            // rows (nu,rho,sigma) of this slab are rows [s0*nbf^2, (s0+sn)*nbf^2) of Q1, whose leading dimension is nbf^3
            if (quarter(h, h->aoQ1.d() + (size_t)s0 * n2, src, dCo, (i64)sn * n2, o, nbf, n3)) return 1;
        }
    }
    // quarter 2: contract nu.  Q2v[(rho,sigma,i), a], Q2o[(rho,sigma,i), j]
    if (h->aoQ2v.ensure((size_t)n2 * o * v * sizeof(double))) return 1;
    if (h->aoQ2o.ensure((size_t)n2 * o * o * sizeof(double))) return 1;
    if (quarter(h, h->aoQ2v.d(), h->aoQ1.d(), dCv, n2 * o, v, nbf)) return 1;
    if (quarter(h, h->aoQ2o.d(), h->aoQ1.d(), dCo, n2 * o, o, nbf)) return 1;
    // quarter 3: contract rho.  Q3vv[(sigma,i,a), b], Q3vo[(sigma,i,a), j], Q3oo[(sigma,i,j), k]
    if (h->aoQ3vv.ensure((size_t)n1 * o * v * v * sizeof(double))) return 1;
    if (h->aoQ3vo.ensure((size_t)n1 * o * v * o * sizeof(double))) return 1;
    if (h->aoQ3oo.ensure((size_t)n1 * o * o * o * sizeof(double))) return 1;
    if (quarter(h, h->aoQ3vv.d(), h->aoQ2v.d(), dCv, n1 * o * v, v, nbf)) return 1;
    if (quarter(h, h->aoQ3vo.d(), h->aoQ2v.d(), dCo, n1 * o * v, o, nbf)) return 1;
    if (quarter(h, h->aoQ3oo.d(), h->aoQ2o.d(), dCo, n1 * o * o, o, nbf)) return 1;
    // quarter 4: contract sigma with Cv -> OVVV[i,a,b,c], OVOV[i,a,j,b], OOOV[i,j,k,a] in the reference's layouts
    if (h->aoOVVV.ensure((size_t)o * v * v * v * sizeof(double))) return 1;
    if (h->aoOVOV.ensure((size_t)o * v * o * v * sizeof(double))) return 1;
    if (h->aoOOOV.ensure((size_t)o * o * o * v * sizeof(double))) return 1;
    if (quarter(h, h->aoOVVV.d(), h->aoQ3vv.d(), dCv, (i64)o * v * v, v, nbf)) return 1;
    if (quarter(h, h->aoOVOV.d(), h->aoQ3vo.d(), dCv, (i64)o * v * o, v, nbf)) return 1;
    if (quarter(h, h->aoOOOV.d(), h->aoQ3oo.d(), dCv, (i64)o * o * o, v, nbf)) return 1;
    const int ao_launches = h->launches;
    if (fpt_upload_conv(h, o, v, T1, T2, h->aoOVVV.d(), h->aoOOOV.d(), h->aoOVOV.d(), fo, fv)) return 1;
    h->launches += ao_launches;
    h->last.h2d_bytes += h2d;
    h->last.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// Sparse AO list (the reference's default conventional container): `nint` symmetry-unique integrals, vals[z] = (mu nu|rho sigma)
// with zero-based indices idx[4z..4z+3] stored as `index_bytes`-wide integers (2: Vector{NTuple{4,Int16}}, 4: Int32).
extern "C" int fpt_upload_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                                    const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                                    const double* fo, const double* fv)
{
    if (!h) return fail("fpt_upload_ao_sparse: NULL handle");
    if (!T1 || !T2 || !Co || !Cv || !fo || !fv || (nint > 0 && (!idx || !vals))) return fail("fpt_upload_ao_sparse: NULL array argument");
    if (nbf < 1 || o < 1 || v < 1 || o + v > nbf || nint < 0) return fail("fpt_upload_ao_sparse: invalid dimensions nbf=%d o=%d v=%d nint=%lld", nbf, o, v, nint);
    if (index_bytes != 2 && index_bytes != 4) return fail("fpt_upload_ao_sparse: index_bytes must be 2 or 4, got %d", index_bytes);
    if (index_bytes == 2 && nbf > 32767) return fail("fpt_upload_ao_sparse: nbf=%d does not fit 16-bit indices", nbf);
    CK(cudaSetDevice(h->dev));
    auto t0 = std::chrono::steady_clock::now();
    const size_t n4 = (size_t)nbf * nbf * nbf * nbf;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (n4 * sizeof(double) > free_b + h->aoDense.cap)
        return fail("fpt_upload_ao_sparse: the dense AO tensor (%.1f GB for nbf=%d) does not fit the device", n4 * 8e-9, nbf);
    if (h->aoDense.ensure(n4 * sizeof(double))) return 1;
    CK(cudaMemsetAsync(h->aoDense.p, 0, n4 * sizeof(double), h->stream));
    double h2d = 0.0;
    if (nint > 0) {
        const bool idx_dev = is_device_ptr(idx);
        const void* didx = idx;
        if (!idx_dev) {
            if (h->sIdx.ensure((size_t)nint * 4 * index_bytes)) return 1;
            CK(cudaMemcpyAsync(h->sIdx.p, idx, (size_t)nint * 4 * index_bytes, cudaMemcpyHostToDevice, h->stream));
            h2d += (double)nint * 4 * index_bytes;
            didx = h->sIdx.p;
        }
        const double* dvals;
        if (stage_in(h, h->sVals, vals, (size_t)nint, &dvals, &h2d)) return 1;
        const int grid = (int)std::min<long long>((nint + 255) / 256, 148LL * 32);
        if (index_bytes == 2)
            expand_sparse_eri_kernel<short><<<grid, 256, 0, h->stream>>>(h->aoDense.d(), (const short*)didx, dvals, nint, nbf);
        else
            expand_sparse_eri_kernel<int><<<grid, 256, 0, h->stream>>>(h->aoDense.d(), (const int*)didx, dvals, nint, nbf);
        CK(cudaGetLastError());
    }
    if (fpt_upload_ao(h, nbf, o, v, T1, T2, h->aoDense.d(), Co, Cv, fo, fv)) return 1;
    h->launches += 1;
    h->last.h2d_bytes += h2d;
    h->last.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

extern "C" int fpt_triples_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                                     const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                                     const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    auto t0 = std::chrono::steady_clock::now();
    if (fpt_upload_ao_sparse(h, nbf, o, v, T1, T2, nint, idx, index_bytes, vals, Co, Cv, fo, fv)) return 1;
    if (fpt_compute(h, 0, -1, Et, nullptr)) return 1;
    h->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (st) *st = h->last;
    return 0;
}

extern "C" int fpt_triples_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                              const double* Co, const double* Cv, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    auto t0 = std::chrono::steady_clock::now();
    if (fpt_upload_ao(h, nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv)) return 1;
    if (fpt_compute(h, 0, -1, Et, nullptr)) return 1;
    h->last.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (st) *st = h->last;
    return 0;
}

extern "C" int fpt_fp64_peak(fpt_handle* h, int variant, double ms_target, double* tflops)
{
    if (!h || !tflops) return fail("fpt_fp64_peak: NULL argument");
    CK(cudaSetDevice(h->dev));
    if (h->out.ensure(sizeof(double))) return 1;
    const int iters = 4096;
    const int grid = h->n_sm * 8;   // 8 CTAs x 8 warps per SM -> 16 warps per SMSP
    // flops per launch
    const double fl = (variant == 0) ? (double)grid * 8 /*warps*/ * iters * 16.0 * 512.0
                                     : (double)grid * 256 /*threads*/ * iters * 16.0 * 2.0;
    auto launch = [&]() {
        if (variant == 0) peak_dmma_kernel<<<grid, 256, 0, h->stream>>>(h->out.d(), iters, 1e-3);
        else peak_dfma_kernel<<<grid, 256, 0, h->stream>>>(h->out.d(), iters, 1e-3);
    };
    launch();
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    // calibrate launch count
    CK(cudaEventRecord(h->ev0, h->stream));
    launch();
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms1 = 0.f;
    CK(cudaEventElapsedTime(&ms1, h->ev0, h->ev1));
    int reps = (int)(ms_target / (ms1 > 1e-3f ? ms1 : 1e-3f));
    if (reps < 1) reps = 1;
    if (reps > 20000) reps = 20000;
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int t = 0; t < reps; t++) launch();
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *tflops = fl * reps / (ms * 1e-3) / 1e12;
    return 0;
}

extern "C" int fpt_set_debug_flags(fpt_handle* h, int flags)
{
    if (!h) return fail("fpt_set_debug_flags: NULL handle");
    h->prob.dbg_flags = flags;
    h->dbg_flags = flags;
    return 0;
}

// Restrict the work list to positions [t_begin, t_end) of the reference's flattened i >= j >= k triplet list (k fastest,
// zero-weight i = j = k entries included, exactly the list the loops of ijk.jl:49,63,83 walk); t_end < 0 = to the end.
extern "C" int fpt_set_triplet_window(fpt_handle* h, long long t_begin, long long t_end)
{
    if (!h) return fail("fpt_set_triplet_window: NULL handle");
    if (!h->loaded) return fail("fpt_set_triplet_window: no problem uploaded");
    Problem& P = h->prob;
    const i64 nfull = (i64)P.o * (P.o + 1) * (P.o + 2) / 6;
    if (t_end < 0 || t_end > nfull) t_end = nfull;
    if (t_begin < 0) t_begin = 0;
    if (t_begin > t_end) t_begin = t_end;
    const i64 u0 = triplets_before(P.o, t_begin), u1 = triplets_before(P.o, t_end);
    P.tw_begin = u0;
    P.tw_count = u1 - u0;
    P.nitems = P.nb * P.tw_count;
    return 0;
}

// 1: block-major (default), 0: triplet-major.  Takes effect for the next compute; keeps the triplet window.
extern "C" int fpt_set_item_order(fpt_handle* h, int order)
{
    if (!h) return fail("fpt_set_item_order: NULL handle");
    if (order != 0 && order != 1) return fail("fpt_set_item_order: order must be 0 or 1, got %d", order);
    h->item_order = order;
    h->prob.order = order;
    return 0;
}

// Part `rank` of `world` of the current work list, as an item range for fpt_compute: contiguous, equal estimated cost.
extern "C" int fpt_shard_items(fpt_handle* h, int rank, int world, long long* item_begin, long long* item_end)
{
    if (!h || !item_begin || !item_end) return fail("fpt_shard_items: NULL argument");
    if (!h->loaded) return fail("fpt_shard_items: no problem uploaded");
    if (world < 1 || rank < 0 || rank >= world) return fail("fpt_shard_items: invalid rank %d of %d", rank, world);
    i64 sb, se;
    shard_range(h, 0, h->prob.nitems, rank, world, &sb, &se);
    *item_begin = sb;
    *item_end = se;
    return 0;
}

extern "C" int fpt_set_kernel_variant(fpt_handle* h, int variant)
{
    if (!h) return fail("fpt_set_kernel_variant: NULL handle");
    if (variant != 1 && variant != 2) return fail("fpt_set_kernel_variant: variant must be 1 or 2, got %d", variant);
    h->kernel_variant = variant;
    return 0;
}

extern "C" int fpt_set_profiling(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_profiling: NULL handle");
    h->profiling = on != 0;
    return 0;
}

// Phase breakdown of the last fpt_compute (cycles summed over CTAs, warp 0's view):
// out[0..5] = setup, zero+prologue, k-loops, RMW epilogues, energy stage, total
extern "C" int fpt_last_profile(fpt_handle* h, double* out6)
{
    if (!h || !out6) return fail("fpt_last_profile: NULL argument");
    if (h->last_grid <= 0 || !h->last_profiled) return fail("fpt_last_profile: the last compute was not profiled (fpt_set_profiling)");
    CK(cudaSetDevice(h->dev));
    std::vector<long long> buf((size_t)h->last_grid * NPROF);
    CK(cudaMemcpy(buf.data(), h->prof.p, buf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int t = 0; t < NPROF; t++) out6[t] = 0.0;
    for (int b = 0; b < h->last_grid; b++)
        for (int t = 0; t < NPROF; t++) out6[t] += (double)buf[(size_t)b * NPROF + t];
    return 0;
}

// DMMA issue study (design aid): sustained TFLOP/s with `ilp` independent accumulators per warp and
// `warps_per_sm` warps on each SM (1 CTA/SM).
extern "C" int fpt_dmma_sweep(fpt_handle* h, int ilp, int warps_per_sm, double* tflops)
{
    if (!h || !tflops) return fail("fpt_dmma_sweep: NULL argument");
    CK(cudaSetDevice(h->dev));
    if (h->out.ensure(sizeof(double))) return 1;
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
    k<<<h->n_sm, threads, smem, h->stream>>>(h->out.d(), iters, 1e-3);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < 5; r++) k<<<h->n_sm, threads, smem, h->stream>>>(h->out.d(), iters, 1e-3);
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *tflops = 5.0 * h->n_sm * warps_per_sm * (double)iters * ilp * 512.0 / (ms * 1e-3) / 1e12;
    return 0;
}
