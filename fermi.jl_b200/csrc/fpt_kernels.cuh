// Hand-written sm_100a kernels of the B200 RCCSD(T) path.
//   K4  prep_*            layout prep (ijk.jl:24-32 replaced by one pass into the Pt/Qt/OV2/T1d layouts)
//   K1+K2 triples_kernel  fused W build (FP64 tensor-core DMMA.8x8x4) + 6-fold permutation + V + denominators +
//                         energy reduction (ijk.jl:108-136); W/V live in shared memory only
//   reduce_partials       fixed-order final sum (ijk.jl:145)
// See fpt_layout.h for the index algebra and DESIGN.md for the data-flow picture.
#pragma once
#include <cuda_runtime.h>
#include "fpt_layout.h"

namespace fpt {

constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
constexpr int QSTAGES = 3;
constexpr int QSTAGE_DOUBLES = CHUNK_GROUPS * TMAX * 2 * KGROUP;       // 1024 doubles = 8 KB
constexpr int WSLOT_DOUBLES = MAX_SLOTS * TMAX * TMAX * TMAX;          // 24576 doubles = 192 KB

struct Ctl {
    i64 cur_item;
    ItemDesc item;
    BlockDesc bd;
    int ngemm;
    GemmDesc gemm[MAX_GEMMS];
    double red[NWARPS];
};

constexpr size_t TRIPLES_SMEM_BYTES = (size_t)(WSLOT_DOUBLES + QSTAGES * QSTAGE_DOUBLES) * sizeof(double) + sizeof(Ctl);

// ---------------------------------------------------------------------------------------------------
// small PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------------------------------------------
// Q stage fill: chunk c of GEMM g  ->  Qsm[stage][gl][zl][s][8]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fill_q_stage(const Problem& P, const Ctl* ctl, double* Qsm, int sp, int nchunks, int tid)
{
    const int g = sp / nchunks;
    if (g < ctl->ngemm) {
        const int c = sp - g * nchunks;
        const GemmDesc& gd = ctl->gemm[g];
        const int g0 = c * CHUNK_GROUPS;
        const int ng = min(CHUNK_GROUPS, P.G - g0);
        const int TZ = gd.TZ;
        double* stage = Qsm + (sp % QSTAGES) * QSTAGE_DOUBLES;
        const int npieces = ng * TZ * 8;  // 16-byte pieces
        for (int pc = tid; pc < npieces; pc += NTHREADS) {
            const int part = pc & 3, s = (pc >> 2) & 1, row = pc >> 3;
            const int gl = row / TZ, zl = row - gl * TZ;
            const double* src = P.Qt + qt_row(P, s ? gd.r : gd.q, s ? gd.q : gd.r, g0 + gl, gd.z0 + zl) + part * 2;
            cp_async16(stage + ((gl * TZ + zl) * 2 + s) * KGROUP + part * 2, src);
        }
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------------------
// One P-stationary GEMM of an item: D[TX*TY x 2*TZ] over all kappa, then RMW into the W slots.
//   MTW = 8-row tiles per warp, NT = 8-column tiles (TZ/4).  Column n of tile ct: zl = 4ct + (n>>1), s = n&1.
// ---------------------------------------------------------------------------------------------------
template <int MTW, int NT>
__device__ __forceinline__ void gemm_body(const Problem& P, const Ctl* ctl, double* Wsm, double* Qsm, int g, int& sp,
                                          int nchunks, int tid, long long* prof)
{
    const long long tg0 = clock64();
    const GemmDesc& gd = ctl->gemm[g];
    const int lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, kk = lane & 3;
    const int TX = gd.TX, TZ = gd.TZ;
    const int rt_total = (TX * gd.TY) >> 3;

    const double* rowp[MTW];
    int xl[MTW], yl[MTW];
    bool valid[MTW];
#pragma unroll
    for (int mt = 0; mt < MTW; mt++) {
        const int rt = warp * MTW + mt;
        valid[mt] = rt < rt_total;
        const int m = valid[mt] ? rt * 8 + r : r;
        yl[mt] = m / TX;
        xl[mt] = m - yl[mt] * TX;
        rowp[mt] = P.Pt + pt_row(P, gd.p, gd.y0 + yl[mt], gd.x0 + xl[mt]) + 2 * kk;
    }

    double acc[MTW][NT][2];
#pragma unroll
    for (int mt = 0; mt < MTW; mt++)
#pragma unroll
        for (int ct = 0; ct < NT; ct++) acc[mt][ct][0] = acc[mt][ct][1] = 0.0;

    double2 a[2][MTW];
#pragma unroll
    for (int mt = 0; mt < MTW; mt++) a[0][mt] = ldg_stream_f64x2(rowp[mt]);

    // B fragment smem offset (doubles) inside a stage for ct = 0, gl = 0
    const int n = lane >> 2;
    const int boff = (((n >> 1)) * 2 + (n & 1)) * KGROUP + 2 * kk;

    for (int c = 0; c < nchunks; c++, sp++) {
        cp_async_wait<1>();
        __syncthreads();
        fill_q_stage(P, ctl, Qsm, sp + 2, nchunks, tid);
        const double* stage = Qsm + (sp % QSTAGES) * QSTAGE_DOUBLES;
        const int g0 = c * CHUNK_GROUPS;
        const int ng = min(CHUNK_GROUPS, P.G - g0);
#pragma unroll
        for (int gl = 0; gl < CHUNK_GROUPS; gl++) {
            if (gl < ng) {
                const int gg = g0 + gl;
                if (gg + 1 < P.G) {
#pragma unroll
                    for (int mt = 0; mt < MTW; mt++) a[(gl + 1) & 1][mt] = ldg_stream_f64x2(rowp[mt] + (gg + 1) * KGROUP);
                }
                double2 b[NT];
#pragma unroll
                for (int ct = 0; ct < NT; ct++)
                    b[ct] = *reinterpret_cast<const double2*>(stage + (gl * TZ + 4 * ct) * 2 * KGROUP + boff);
#pragma unroll
                for (int mt = 0; mt < MTW; mt++)
#pragma unroll
                    for (int ct = 0; ct < NT; ct++) dmma884(acc[mt][ct][0], acc[mt][ct][1], a[gl & 1][mt].x, b[ct].x);
#pragma unroll
                for (int mt = 0; mt < MTW; mt++)
#pragma unroll
                    for (int ct = 0; ct < NT; ct++) dmma884(acc[mt][ct][0], acc[mt][ct][1], a[gl & 1][mt].y, b[ct].y);
            }
        }
    }

    const long long tg1 = clock64();
    // RMW epilogue: D element e of (mt,ct): row (xl,yl), zl = 4ct + kk, column set s = e.
    // When X and Z are the same tile, D(s=0)[x=u,z=w] and D(s=1)[x=w,z=u] of *different* warps hit the same W
    // element, so the two column sets are separated by a barrier.
#pragma unroll
    for (int e = 0; e < 2; e++) {
        if (e == 1 && gd.diag_xz) __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MTW; mt++) {
            if (valid[mt]) {
#pragma unroll
                for (int ct = 0; ct < NT; ct++) {
                    const int off = gemm_dest(gd, e, xl[mt], yl[mt], 4 * ct + kk);
                    Wsm[off] += acc[mt][ct][e];
                }
            }
        }
    }
    prof[2] += tg1 - tg0;
    prof[3] += clock64() - tg1;
}

template <int MTW>
__device__ __forceinline__ void gemm_dispatch_nt(const Problem& P, const Ctl* ctl, double* Wsm, double* Qsm, int g, int& sp,
                                                 int nchunks, int tid, int nt, long long* prof)
{
    switch (nt) {
    case 1: gemm_body<MTW, 1>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, prof); break;
    case 2: gemm_body<MTW, 2>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, prof); break;
    case 3: gemm_body<MTW, 3>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, prof); break;
    default: gemm_body<MTW, 4>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, prof); break;
    }
}

// ---------------------------------------------------------------------------------------------------
// Energy stage for the block held in Wsm: ijk.jl:116 (V) and :120-136 (a>=b>=c loop, weights, denominators)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_energy(const Problem& P, const Ctl* ctl, const double* Wsm, int tid)
{
    const BlockDesc& bd = ctl->bd;
    const int i = ctl->item.i, j = ctl->item.j, k = ctl->item.k;
    const int npts = bd.slot_elems;
    double e = 0.0;
    for (int pt = tid; pt < npts; pt += NTHREADS) e += block_point_energy(P, bd, i, j, k, Wsm, pt);
    return e;
}

// ---------------------------------------------------------------------------------------------------
// The fused persistent kernel.  grid = #SMs (1 CTA/SM: ~220 KB smem), block = 256.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
triples_kernel(Problem P, i64 item_begin, i64 item_end, unsigned long long* counter, double* partials, long long* prof_out)
{
    long long prof[6] = {0, 0, 0, 0, 0, 0};   // cycles: setup, zero+prologue, k-loops, RMW, energy, total (warp 0's view)
    const long long tk0 = clock64();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Wsm = reinterpret_cast<double*>(smem_raw);
    double* Qsm = Wsm + WSLOT_DOUBLES;
    Ctl* ctl = reinterpret_cast<Ctl*>(Qsm + QSTAGES * QSTAGE_DOUBLES);
    const int tid = threadIdx.x;
    const int nchunks = (P.G + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
    double esum = 0.0;

    for (;;) {
        const long long ts0 = clock64();
        __syncthreads();
        if (tid == 0) {
            const i64 it = item_begin + (i64)atomicAdd(counter, 1ULL);
            ctl->cur_item = it;
            if (it < item_end) {
                item_decode(P, it, ctl->item);
                make_block(ctl->item.A, ctl->item.B, ctl->item.C, P.vp, ctl->bd);
                ctl->ngemm = make_gemms(ctl->bd, ctl->item.i, ctl->item.j, ctl->item.k, ctl->gemm);
            }
        }
        __syncthreads();
        if (ctl->cur_item >= item_end) break;
        const long long ts1 = clock64();
        prof[0] += ts1 - ts0;

        {   // zero the live W slots
            const int nz = ctl->bd.nslot * ctl->bd.slot_elems;   // multiple of 64
            double2* w2 = reinterpret_cast<double2*>(Wsm);
            for (int idx = tid; idx < nz / 2; idx += NTHREADS) w2[idx] = make_double2(0.0, 0.0);
        }
        int sp = 0;
        fill_q_stage(P, ctl, Qsm, 0, nchunks, tid);
        fill_q_stage(P, ctl, Qsm, 1, nchunks, tid);
        const int ngemm = ctl->ngemm;
        prof[1] += clock64() - ts1;
        for (int g = 0; g < ngemm; g++) {
            const GemmDesc& gd = ctl->gemm[g];
            const int rt_total = (gd.TX * gd.TY) >> 3;
            const int mtw = (rt_total + NWARPS - 1) / NWARPS;
            const int nt = gd.TZ >> 2;
            switch (mtw) {
            case 1: gemm_dispatch_nt<1>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, nt, prof); break;
            case 2: gemm_dispatch_nt<2>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, nt, prof); break;
            case 3: gemm_dispatch_nt<3>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, nt, prof); break;
            default: gemm_dispatch_nt<4>(P, ctl, Wsm, Qsm, g, sp, nchunks, tid, nt, prof); break;
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        const long long te0 = clock64();
        esum += block_energy(P, ctl, Wsm, tid);
        prof[4] += clock64() - te0;
    }
    prof[5] = clock64() - tk0;
    if (prof_out && tid == 0)
        for (int t = 0; t < 6; t++) prof_out[blockIdx.x * 6 + t] = prof[t];

    // block reduction (warp shuffle, then one thread sums the 8 warp partials in fixed order)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, off);
    if ((tid & 31) == 0) ctl->red[tid >> 5] = esum;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < NWARPS; w++) s += ctl->red[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void reduce_partials(const double* partials, int n, double* out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < n; t++) s += partials[t];
        out[0] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// K4: layout prep.  Sources are the reference's column-major arrays (first index fastest).
// ---------------------------------------------------------------------------------------------------
// Pt[p][y][x][d] = OVVV[p,y,x,d] for d in [d0, d0+dn); `src` points at OVVV[:,:,:,d0].
__global__ void prep_pt_particle(Problem P, double* Pt, const double* __restrict__ src, int d0, int dn)
{
    __shared__ double tile[32][33];
    const int o = P.o, v = P.v;
    const i64 ov = (i64)o * v;
    const int x = blockIdx.z;
    const i64 py0 = (i64)blockIdx.x * 32;
    const int dd0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int kk = 0; kk < 32; kk += 8) {
        const int dd = dd0 + ty + kk;
        const i64 py = py0 + tx;
        if (dd < dn && py < ov) tile[ty + kk][tx] = src[py + ov * ((i64)x + (i64)v * dd)];
    }
    __syncthreads();
    for (int kk = 0; kk < 32; kk += 8) {
        const i64 py = py0 + ty + kk;
        const int dd = dd0 + tx;
        if (dd < dn && py < ov) {
            const int p = (int)(py % o), y = (int)(py / o);
            Pt[pt_row(P, p, y, x) + d0 + dd] = tile[tx][ty + kk];
        }
    }
}

// Pt[p][y][x][v+l] = -T2[p,l,y,x]
__global__ void prep_pt_hole(Problem P, double* Pt, const double* __restrict__ T2)
{
    const int o = P.o, v = P.v;
    const i64 n = (i64)o * v * v * o;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int l = (int)(idx % o);
        i64 t = idx / o;
        const int x = (int)(t % v); t /= v;
        const int y = (int)(t % v);
        const int p = (int)(t / v);
        Pt[pt_row(P, p, y, x) + v + l] = -T2[p + (i64)o * (l + (i64)o * (y + (i64)v * x))];
    }
}

// Qt[(q,r)][g][z][kk8]: kappa<v: T2[r,q,z,kappa]; v<=kappa<v+o: OOOV[kappa-v,q,r,z]; else 0
__global__ void prep_qt(Problem P, double* Qt, const double* __restrict__ T2, const double* __restrict__ OOOV)
{
    const int o = P.o, v = P.v;
    const i64 n = (i64)o * o * P.G * P.vp * KGROUP;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int k8 = (int)(idx % KGROUP);
        i64 t = idx / KGROUP;
        const int z = (int)(t % P.vp); t /= P.vp;
        const int g = (int)(t % P.G); t /= P.G;
        const int r = (int)(t % o);
        const int q = (int)(t / o);
        const int kappa = g * KGROUP + k8;
        double val = 0.0;
        if (z < v) {
            if (kappa < v) val = T2[r + (i64)o * (q + (i64)o * (z + (i64)v * kappa))];
            else if (kappa < v + o) val = OOOV[(kappa - v) + (i64)o * (q + (i64)o * (r + (i64)o * z))];
        }
        Qt[idx] = val;
    }
}

// OV2[(q,r)][y][z] = OVOV[q,y,r,z]
__global__ void prep_ov2(Problem P, double* OV2, const double* __restrict__ OVOV)
{
    const int o = P.o, v = P.v;
    const i64 n = (i64)o * o * v * v;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int z = (int)(idx % v);
        i64 t = idx / v;
        const int y = (int)(t % v); t /= v;
        const int r = (int)(t % o);
        const int q = (int)(t / o);
        OV2[idx] = OVOV[q + (i64)o * (y + (i64)v * (r + (i64)o * z))];
    }
}

// T1d[p][x] = T1[p,x]
__global__ void prep_t1(Problem P, double* T1d, const double* __restrict__ T1)
{
    const int o = P.o, v = P.v;
    const int n = o * v;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int x = idx % v, p = idx / v;
        T1d[idx] = T1[p + (i64)o * x];
    }
}

// ---------------------------------------------------------------------------------------------------
// FP64 pipe calibration (roofline denominator): register-resident DMMA.8x8x4 / DFMA streams
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x * 1e-9, a1 = a0 * 0.5, a2 = a0 * 0.25, a3 = a0 * 0.125;
    double b0 = 1.0 + a0, b1 = 1.0 - a0, b2 = 0.5 + a0, b3 = 0.5 - a0;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < iters; it++) {
        const double av[4] = {a0, a1, a2, a3}, bv[4] = {b0, b1, b2, b3};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i], bv[j]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;   // keep the work alive
}

// DMMA issue study: ILP independent accumulators per warp, launched with 1 CTA/SM and a chosen warp count
template <int ILP>
__global__ void dmma_ilp_kernel(double* out, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - a;
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i][0] = acc[i][1] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - a * 1e-3;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = i * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], b, a);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace fpt

// ---------------------------------------------------------------------------------------------------
// K3: density-fitted assembly.  C(m,n) = sum_Q A[Q + naux*rowA(m)] * B[Q + naux*rowB(n)] on DMMA.8x8x4, written
// straight into the device layouts (the o*v^3 tensor never exists on the host -- the reference materialises it in
// DFERI.jl:156-180).  MODE 0: Pt particle part from BOV,BVV.  MODE 1: Qt hole part (OOOV) from BOO,BOV.
// MODE 2: OV2 (OVOV) from BOV,BOV.  CTA = 4 warps, 64x64 tile, fragments loaded straight from global (L1-shared).
// ---------------------------------------------------------------------------------------------------
namespace fpt {

template <int MODE>
__device__ __forceinline__ void df_store(const Problem& P, double* out, int m, int n, double val)
{
    const int o = P.o, v = P.v;
    if (MODE == 0) {
        const int p = m % o, y = m / o, d = n % v, x = n / v;
        out[pt_row(P, p, y, x) + d] = val;
    } else if (MODE == 1) {
        const int l = m % o, q = m / o, r = n % o, z = n / o;
        const int kappa = v + l;
        out[qt_row(P, q, r, kappa >> 3, z) + (kappa & 7)] = val;
    } else {
        const int q = m % o, y = m / o, r = n % o, z = n / o;
        out[(((i64)q * o + r) * v + y) * v + z] = val;
    }
}

template <int MODE>
__global__ void __launch_bounds__(128) df_gemm_kernel(Problem P, double* out, const double* __restrict__ A,
                                                      const double* __restrict__ B, int M, int N, int naux)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = lane >> 2, kk = lane & 3;
    const int m0 = blockIdx.x * 64 + (warp >> 1) * 32;
    const int n0 = blockIdx.y * 64 + (warp & 1) * 32;
    const double* ap[4];
    const double* bp[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int m = m0 + 8 * t + r; if (m >= M) m = M - 1;
        ap[t] = A + (i64)naux * m;
        int n = n0 + 8 * t + r; if (n >= N) n = N - 1;
        int rowb = n;
        if (MODE == 0) { const int d = n % P.v, x = n / P.v; rowb = x + P.v * d; }
        bp[t] = B + (i64)naux * rowb;
    }
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < naux; k0 += 4) {
        const int k = k0 + kk;
        const bool ok = k < naux;
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; t++) { a[t] = ok ? __ldg(ap[t] + k) : 0.0; b[t] = ok ? __ldg(bp[t] + k) : 0.0; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int m = m0 + 8 * i + r, n = n0 + 8 * j + 2 * kk + e;
                if (m < M && n < N) df_store<MODE>(P, out, m, n, acc[i][j][e]);
            }
}

}  // namespace fpt
