// Host-side state of libfermi_pt_b200.so: error reporting, device buffers, the per-GPU state (`Dev`), the handle, and the NCCL
// entry points (resolved at run time, only handles that span more than one GPU need them).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fermi_pt_b200.h"
#include "fpt_layout.h"
#include "fpt_stage.h"

namespace fpt {

inline std::string& err_text()
{
    static thread_local std::string s;
    return s;
}
inline int fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err_text() = buf;
    return 1;
}
#define CK(call)                                                                                                             \
    do {                                                                                                                     \
        cudaError_t e_ = (call);                                                                                             \
        if (e_ != cudaSuccess) return fpt::fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

// every entry point leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)   // the owning device must be current
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

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    bool ok = false;
    std::string why;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
inline NcclApi& nccl_api()
{
    static NcclApi api;
    return api;
}
inline int nccl_load()
{
    static std::once_flag once;
    NcclApi& g = nccl_api();
    std::call_once(once, [&g] {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { g.why = std::string("cannot load libnccl.so.2 (") + dlerror() + ")"; return; }
#define FPT_SYM(field, name)                                                        \
    g.field = reinterpret_cast<decltype(g.field)>(dlsym(lib, name));                \
    if (!g.field) { g.why = std::string("libnccl lacks ") + name; return; }
        FPT_SYM(CommInitAll, "ncclCommInitAll");
        FPT_SYM(CommInitRank, "ncclCommInitRank");
        FPT_SYM(GetUniqueId, "ncclGetUniqueId");
        FPT_SYM(CommDestroy, "ncclCommDestroy");
        FPT_SYM(GroupStart, "ncclGroupStart");
        FPT_SYM(GroupEnd, "ncclGroupEnd");
        FPT_SYM(Broadcast, "ncclBroadcast");
        FPT_SYM(AllGather, "ncclAllGather");
        FPT_SYM(AllReduce, "ncclAllReduce");
        FPT_SYM(GetErrorString, "ncclGetErrorString");
#undef FPT_SYM
        g.lib = lib;
        g.ok = true;
    });
    if (!g.ok) return fail("multi-GPU handle: %s", g.why.c_str());
    return 0;
}
#define NCK(call)                                                                                                                        \
    do {                                                                                                                                 \
        ncclResult_t r_ = (call);                                                                                                        \
        if (r_ != ncclSuccess) return fpt::fail("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, fpt::nccl_api().GetErrorString(r_)); \
    } while (0)

// ---- one GPU of a handle -----------------------------------------------------------------------------------------------
constexpr int MAX_PHASES = 4;   // kernels of one split call (triples_conv in fpt_api.cu)
constexpr int MAX_WORLD = 16;    // GPUs of one communicator (fpt_create refuses more)
constexpr int OUT_DOUBLES = 1 + MAX_PHASES * MAX_WORLD;   // E(T) + one time slot per (phase, GPU), see ShardCal
constexpr int NTL = 6;   // timeline events: upload begin, last H2D done, operands ready, kernel begin, kernel end, result ready
struct Dev {
    int dev = 0;
    int idx = 0;          // position in fpt_handle::devs
    int grank = 0;        // rank in the NCCL communicator (= global shard number)
    int n_sm = 0;
    size_t total_mem = 0;     // device memory (bytes)
    cudaStream_t stream = nullptr;   // kernels + collectives
    cudaStream_t copy = nullptr;     // host -> device DMAs
    cudaEvent_t ev0[MAX_PHASES] = {}, ev1[MAX_PHASES] = {};   // around the fused kernel of each phase of a call (one phase unless split)
    cudaEvent_t ev_copy = nullptr, ev_start = nullptr;   // copy -> stream / stream -> copy hand-offs
    cudaEvent_t ev_free[2] = {nullptr, nullptr};         // OVVV chunk buffer c&1 has been consumed by its prep kernel
    cudaEvent_t tl[NTL] = {};
    ncclComm_t comm = nullptr;
    // resident operands
    DevBuf Pt, Qt, OV2, T1d, fo, fv, partials, counter, out, prof, blocktab;
    // raw inputs (staging)
    DevBuf sT1, sT2, sOOOV, sOVOV, sChunk[2], sPhase[MAX_PHASES], sTri, sTri2, sBOO, sBOV, sBVV;
    // AO -> MO route: coefficient blocks, quarter-transformed intermediates, the MO blocks the (T) path consumes
    DevBuf sCo, sCv, aoDense, sIdx, sVals, aoQ1, aoQ2v, aoQ2o, aoQ3vv, aoQ3vo, aoQ3oo, aoOVVV, aoOOOV, aoOVOV, aoFlag;
    // CCSD ladder / MP2: tau^T, the (vv|vv) slabs of one group of a, newT2 on the device, (ia|jb)
    DevBuf xTau, xSlab, xNew, xOVOV;
    static constexpr int NF32 = 7;
    DevBuf f32in[NF32], f32wide[NF32];   // Float32 callers: the arrays as they arrive and their widened images
    DevBuf ringtab;           // slab-ring mode of the DF route: triplet lists and slot maps of all block triples
    const double* cur_T2 = nullptr;   // T2 of the current problem on this GPU (the ring re-assembles Pt's hole part from it)
    int pt_slabs = 0;         // occupied slabs Pt has room for (o, or 3 ob in ring mode)
    Problem prob{};
    int tab_vp = -1;          // the block table on this device was built for this padded virtual dimension
    // Pt holds zeros in its padding (x,y >= v, kappa >= v+o) for this shape: a new upload of the same shape may skip the memset
    int clean_o = -1, clean_v = -1, clean_slabs = 0;
    void* clean_ptr = nullptr;
    int last_grid = 0;
    double last_ms[MAX_PHASES] = {};     // kernel time of this GPU's shard in the last call, per phase ...
    int last_gen[MAX_PHASES] = {};       // ... and the generation of boundary fractions it was measured with (0: none)
    i64 shard_b = 0, shard_e = 0;
};

}  // namespace fpt

// Adaptive balance of the static shards of a multi-GPU handle.  The split is by *estimated* cost (block_cost); at C4 on 8 GPUs the
// shards' kernel times differ by +-2 %.  Repeated calls of one shape (the 6 N_atoms calls of a finite-difference gradient) correct it:
// every GPU's kernel time of call n travels with the scalar all-reduce of call n+1 (one slot per phase and GPU in the reduced
// vector), so after two calls with the same boundaries every process holds all times measured with them and moves the boundaries to
// where the measured time density says equal times lie (damped).  All processes see the same vector and run the same arithmetic,
// so the new boundaries agree everywhere without further communication.
struct ShardCal {
    int o = -1, v = -1, world = 0, order = -1;
    fpt::i64 tw_begin = -1, tw_count = -1, b = -1, e = -1;      // what was split
    int gen = 0;                                                // generation of `frac` (0: uniform, nothing measured)
    std::vector<double> frac;                                   // world + 1 boundary fractions in use
    bool pending = false;                                       // all times of this generation are known: move the boundaries at the next launch
    std::vector<double> ms;                                     // the times (per GPU)
};

struct fpt_handle {
    std::vector<fpt::Dev*> devs;   // the GPUs this process drives
    int world = 1;                 // GPUs in the communicator (== devs.size() unless created with fpt_create_rank)
    bool rank_mode = false;        // one process per GPU: uploads and computes are collective calls over `world` processes
    fpt::StagePool pool;           // host threads + pinned bounce slots for pageable inputs
    double* res_pinned = nullptr;  // E(T) (and the time slots of the adaptive balance) land here: OUT_DOUBLES doubles
    ShardCal cal[fpt::MAX_PHASES];
    int adaptive = 1;              // fpt_set_adaptive_shards (must be the same on every rank of a communicator: it sizes the all-reduce)
    int gen_counter = 0;           // source of ShardCal::gen
    bool ring_call = false;        // the evaluation in flight / last finished used the DF slab ring (no time slots)
    // problem (identical on every GPU)
    int o = 0, v = 0;
    std::vector<fpt::BlockTabEntry> tab;
    int tab_vp = -1;
    std::vector<double> block_cost;
    fpt::i64 tw_begin = 0, tw_count = 0, nitems = 0;
    int item_order = 1;       // 1: block-major (default), 0: triplet-major (see Problem::order)
    int dbg_flags = 0;
    int deterministic = 0;    // fpt_set_deterministic: static item -> CTA deal, E(T) bitwise reproducible from run to run
    int df_ring = 0;          // fpt_set_df_ring: -1 never, 0 automatic, n >= 1 slab ring with occupied blocks of n
    int sym_inputs = 1;       // pageable host inputs cross PCIe as their symmetry-unique halves (fpt_set_symmetric_inputs)
    bool profiling = false, last_profiled = false;
    int kernel_variant = 1;
    bool loaded = false;
    bool pending = false;     // an asynchronous call is in flight (fpt_wait has to collect it)
    int nphase = 1;           // kernels per GPU of the evaluation in flight / last finished (> 1: split call, see triples_conv)
    fpt::i64 pend_items = 0;
    fpt_stats last{};
    int launches = 0;
    double h2d = 0.0, stage_host_ms = 0.0;
    double timeline[8] = {};
};
