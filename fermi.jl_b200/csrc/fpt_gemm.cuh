// K3 / K5: FP64 tensor-core GEMM of two operands that both have the contracted index contiguous ("TN" form),
//
//     C(m, n) = sum_q  A[q + K * rowA(m)] * B[q + K * rowB(n)],      m < M, n < N,
//
// used for (K3) the density-fitted assembly of the (T) operands from BOO / BOV / BVV (DFERI.jl:88-180 builds the same blocks
// on the CPU with @tensoropt) and (K5) the four quarter transformations AO -> MO (Chonky.jl:28-114).  The epilogue writes the
// result straight into the layout its consumer wants (Pt / Qt / OV2 of the fused kernel, or a column-major matrix).
//
// Persistent CTAs (one per SM) of 16 warps, tile 128 x BN (BN = 128, or 32 for the transforms onto the occupied space), K in chunks of 32.  Operand
// tiles are staged in shared memory by cp.async (SASS LDGSTS) through a 3-stage ring shared by all warps -- every operand
// element is fetched from L2 once per CTA -- in rows of 34 doubles (16 bytes of skew per row: the 32-byte fragment loads of
// a quarter-warp then cover all 32 banks).  A lane's fragment for one 16-wide kappa group is 4 consecutive kappa (one
// 32-byte shared-memory load) used by four DMMA.8x8x4: the order of the contraction index inside a group is free as long as
// A and B agree.  sm_100a has no f64 tcgen05 kind: DMMA is the FP64 tensor path.
// K and the row maps are arbitrary (naux = 37, nbf = 9 ...): rows are only 8-byte aligned then, so the copies are 8 bytes
// wide unless K is even and the bases are 16-byte aligned (ALIGN16); the K tail is zero-filled by cp.async's src-size form.
#pragma once
#include <cuda_runtime.h>
#include "fpt_layout.h"
#include "fpt_ptx.cuh"

namespace fpt {

// row(m) = base + (m % n0) * s0 + (m / n0) * s1
struct RowMap {
    i64 base, s0, s1;
    int n0;
};
__host__ __device__ inline RowMap rowmap_identity() { return RowMap{0, 1, 0, 0x7fffffff}; }
__device__ __forceinline__ i64 rowmap_apply(const RowMap& r, i64 m)
{
    if (r.n0 == 0x7fffffff) return r.base + m * r.s0;
    return r.base + (m % r.n0) * r.s0 + (m / r.n0) * r.s1;
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_KC = 32;                 // kappa per stage
constexpr int GEMM_LDS = GEMM_KC + 2;       // doubles per staged row (272 B: 16 B skew mod 128 B)
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 512;
constexpr int GEMM_ROWOFF_RING = 4;         // row-offset tables kept: tiles j-1 .. j+2 of a CTA's tile sequence
template <int BN>
constexpr size_t gemm_smem_bytes() { return (size_t)GEMM_STAGES * (GEMM_BM + BN) * GEMM_LDS * sizeof(double) + (size_t)GEMM_ROWOFF_RING * (GEMM_BM + BN) * sizeof(i64); }

__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- epilogues ------------------------------------------------------------------------------------------------------
// EPI_COLMAJOR  C[m + ldc*n]                                                      (quarter transforms)
// EPI_PT        m = y + v*pl, n = d + v*x  ->  Pt[p0+pl][y][x][d]                 (OVVV[p,y,x,d] = sum_Q BOV[Q,p,y] BVV[Q,x,d], DFERI.jl:156-180)
// EPI_QT_HOLE   m = l + o*q,  n = r + o*z  ->  Qt[(q,r)][g][z][kappa = v + l]     (OOOV[l,q,r,z] = sum_Q BOO[Q,l,q] BOV[Q,r,z], DFERI.jl:88-112)
// EPI_OV2       m = q + o*y,  n = r + o*z  ->  OV2[(q,r)] tile (y,z)              (OVOV[q,y,r,z] = sum_Q BOV[Q,q,y] BOV[Q,r,z], DFERI.jl:139-154)
// EPI_LADDER_SLAB  m = c + v*al, n = d + v*b  ->  X[c + v*d + v^2*b + v^3*al]    ((ca|db) = sum_Q BVV[Q,c,a] BVV[Q,d,b] for a = a0 + al,
//                                                                               RCCSDHelper.jl:216; X_a is the B operand of the next GEMM)
// EPI_LADDER_OUT   m = i + o*j,  n = b + v*al  ->  C[m + o^2*((a0+al) + v*b)] += val   (newT2[:,:,a,:] += tau . X_a, RCCSDHelper.jl:217-218)
enum { EPI_COLMAJOR = 0, EPI_PT = 1, EPI_QT_HOLE = 2, EPI_OV2 = 3, EPI_LADDER_SLAB = 4, EPI_LADDER_OUT = 5 };
struct GemmOut {
    double* C;
    i64 ldc;       // EPI_COLMAJOR
    Problem P;     // the layout epilogues
    int p0;        // EPI_PT: first occupied index of this launch (the assembly is sharded over p across GPUs)
    int lv, lo2, la0;   // EPI_LADDER_*: v, o^2, first a of this group
    const int* pslot;   // EPI_PT: slab position of every occupied index (slab ring of the DF route), nullptr = identity
};
template <int EPI>
__device__ __forceinline__ void gemm_store(const GemmOut& out, i64 m, int n, double val)
{
    const Problem& P = out.P;
    if (EPI == EPI_COLMAJOR) {
        out.C[m + out.ldc * n] = val;
    } else if (EPI == EPI_PT) {
        const int y = (int)(m % P.v), pl = (int)(m / P.v), d = n % P.v, x = n / P.v;
        out.C[pt_row_slot(P, out.pslot, out.p0 + pl, y, x) + d] = val;
    } else if (EPI == EPI_QT_HOLE) {
        const int l = (int)(m % P.o), q = (int)(m / P.o), r = n % P.o, z = n / P.o;
        const int kappa = P.v + l;
        out.C[qt_row(P, q, r, kappa / KGROUP, z) + (kappa % KGROUP)] = val;
    } else if (EPI == EPI_LADDER_SLAB) {
        const i64 v = out.lv;
        out.C[(m % v) + v * (i64)n + v * v * v * (m / v)] = val;
    } else if (EPI == EPI_LADDER_OUT) {
        const int b = n % out.lv, al = n / out.lv;
        out.C[m + (i64)out.lo2 * ((out.la0 + al) + (i64)out.lv * b)] += val;
    } else {
        const int q = (int)(m % P.o), y = (int)(m / P.o), r = n % P.o, z = n / P.o;
        out.C[ov2_idx(P, q, r, y, z)] = val;
    }
}

// MT x NT DMMA tiles per warp; warps laid out WM x WN with WM*MT*8 = 128, WN*NT*8 = BN.
// Persistent: CTA b works on output tiles b, b + grid, ... (m fastest, so concurrently running CTAs share their B rows in L2),
// and the (tile, kappa-chunk) steps of all its tiles form ONE software pipeline: while a tile's last chunks are multiplied and
// its result is stored, the first chunks of the next tile are already in flight.  With K = naux or nbf a tile has only 5-16
// chunks, so a per-tile prologue would idle the tensor pipe for a fifth of the time.
template <int BN, int EPI, bool ALIGN16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const double* __restrict__ A, RowMap mapA, const double* __restrict__ B, RowMap mapB, i64 M, int N, int K, GemmOut out)
{
    constexpr int NT = BN == 128 ? 4 : 2, MT = BN == 128 ? 4 : 2;
    constexpr int WN = BN / (8 * NT), WM = GEMM_BM / (8 * MT);
    static_assert(WM * WN == GEMM_THREADS / 32, "warp grid must use all 16 warps");
    constexpr int ROWS = GEMM_BM + BN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    i64* rowoff = reinterpret_cast<i64*>(tiles + (size_t)GEMM_STAGES * ROWS * GEMM_LDS);   // [GEMM_ROWOFF_RING][ROWS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
    const i64 ntiles = tiles_m * ((N + BN - 1) / BN);
    const i64 nmine = ntiles > blockIdx.x ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int nk = (K + GEMM_KC - 1) / GEMM_KC;

    // element offsets of the staged rows of my j-th tile (A rows first, then B rows); rows past the edge are clamped (their
    // results are not stored).  Ring of 4: tile j+3's offsets are written when tile j is finished.
    auto fill_rowoff = [&](i64 j) {
        const i64 t = blockIdx.x + j * gridDim.x;
        const i64 m0 = (t % tiles_m) * GEMM_BM;
        const i64 n0 = (t / tiles_m) * BN;
        i64* ro = rowoff + (j & (GEMM_ROWOFF_RING - 1)) * ROWS;
        for (int r = tid; r < ROWS; r += GEMM_THREADS) {
            if (r < GEMM_BM) {
                i64 m = m0 + r; if (m >= M) m = M - 1;
                ro[r] = rowmap_apply(mapA, m) * (i64)K;
            } else {
                i64 n = n0 + (r - GEMM_BM); if (n >= N) n = N - 1;
                ro[r] = rowmap_apply(mapB, n) * (i64)K;
            }
        }
    };
    // the load pipeline runs GEMM_STAGES-1 steps ahead of the multiply pipeline and has its own (tile, chunk, stage) counters
    i64 lj = 0;
    int lkt = 0, lstage = 0;
    auto load_next = [&]() {
        if (lj < nmine) {
            const i64* ro = rowoff + (lj & (GEMM_ROWOFF_RING - 1)) * ROWS;
            double* st = tiles + (size_t)lstage * ROWS * GEMM_LDS;
            const int k0 = lkt * GEMM_KC;
            if (ALIGN16) {
                constexpr int CPR = GEMM_KC / 2;   // 16-byte chunks per row
#pragma unroll
                for (int c = tid; c < ROWS * CPR; c += GEMM_THREADS) {
                    const int r = c / CPR, kc = (c % CPR) * 2;
                    const double* base = (r < GEMM_BM) ? A : B;
                    int nb = (K - (k0 + kc)) * 8;
                    nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
                    const i64 koff = nb > 0 ? (k0 + kc) : 0;
                    cp_async16(st + r * GEMM_LDS + kc, base + ro[r] + koff, nb);
                }
            } else {
#pragma unroll 4
                for (int c = tid; c < ROWS * GEMM_KC; c += GEMM_THREADS) {
                    const int r = c / GEMM_KC, kc = c % GEMM_KC;
                    const double* base = (r < GEMM_BM) ? A : B;
                    const bool ok = k0 + kc < K;
                    cp_async8(st + r * GEMM_LDS + kc, base + ro[r] + (ok ? k0 + kc : 0), ok ? 8 : 0);
                }
            }
            if (++lkt == nk) { lkt = 0; lj++; }
            if (++lstage == GEMM_STAGES) lstage = 0;
        }
        cp_async_commit();
    };

    for (i64 j = 0; j < GEMM_ROWOFF_RING - 1 && j < nmine; j++) fill_rowoff(j);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; s++) load_next();
    const int wm = warp / WN, wn = warp % WN;
    const int r8 = lane >> 2, kk = lane & 3;
    const double* afrag = tiles + (wm * MT * 8 + r8) * GEMM_LDS + 4 * kk;                 // + stage, + 8*i rows, + 16*g + 2*hh
    const double* bfrag = tiles + (GEMM_BM + wn * NT * 8 + r8) * GEMM_LDS + 4 * kk;
    int cstage = 0;
    for (i64 j = 0; j < nmine; j++) {
        double acc[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int jj = 0; jj < NT; jj++) acc[i][jj][0] = acc[i][jj][1] = 0.0;
        for (int kt = 0; kt < nk; kt++) {
            cp_async_wait<GEMM_STAGES - 2>();
            __syncthreads();   // this step has landed for every thread; the stage of the previous step is no longer read by anyone
            load_next();
            const int so = cstage * ROWS * GEMM_LDS;
#pragma unroll
            for (int g = 0; g < GEMM_KC / 16; g++) {
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    double2 a[MT], b[NT];
#pragma unroll
                    for (int i = 0; i < MT; i++)
                        a[i] = *reinterpret_cast<const double2*>(afrag + so + 8 * i * GEMM_LDS + 16 * g + 2 * hh);
#pragma unroll
                    for (int jj = 0; jj < NT; jj++)
                        b[jj] = *reinterpret_cast<const double2*>(bfrag + so + 8 * jj * GEMM_LDS + 16 * g + 2 * hh);
#pragma unroll
                    for (int i = 0; i < MT; i++)
#pragma unroll
                        for (int jj = 0; jj < NT; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i].x, b[jj].x);
#pragma unroll
                    for (int i = 0; i < MT; i++)
#pragma unroll
                        for (int jj = 0; jj < NT; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i].y, b[jj].y);
                }
            }
            if (++cstage == GEMM_STAGES) cstage = 0;
        }
        // tile finished: store it (the next tile's first chunks are already in flight)
        const i64 t = blockIdx.x + j * gridDim.x;
        const i64 m0 = (t % tiles_m) * GEMM_BM;
        const int n0 = (int)((t / tiles_m) * BN);
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int jj = 0; jj < NT; jj++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const i64 m = m0 + wm * MT * 8 + 8 * i + r8;
                    const int n = n0 + wn * NT * 8 + 8 * jj + 2 * kk + e;
                    if (m < M && n < N) gemm_store<EPI>(out, m, n, acc[i][jj][e]);
                }
        // the offsets of tile j+3 replace those of tile j-1; they are first read two barriers from now at the earliest
        if (j + GEMM_ROWOFF_RING - 1 < nmine) fill_rowoff(j + GEMM_ROWOFF_RING - 1);
    }
    cp_async_wait<0>();
}

// host-side launch (stream-ordered); returns the CUDA error of the launch
template <int EPI>
inline cudaError_t gemm_tn_launch(cudaStream_t stream, const double* A, RowMap mapA, const double* B, RowMap mapB, i64 M, int N, int K,
                                  const GemmOut& out, int n_sm = 148)
{
    if (M <= 0 || N <= 0) return cudaSuccess;
    const bool al = (K % 2 == 0) && (((uintptr_t)A | (uintptr_t)B) % 16 == 0);
    const bool narrow = N <= 48;
    const i64 tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (narrow ? (N + 31) / 32 : (N + 127) / 128);
    const unsigned grid = (unsigned)(tiles < n_sm ? tiles : n_sm);
    if (narrow) {
        if (al) gemm_tn_kernel<32, EPI, true><<<grid, GEMM_THREADS, gemm_smem_bytes<32>(), stream>>>(A, mapA, B, mapB, M, N, K, out);
        else gemm_tn_kernel<32, EPI, false><<<grid, GEMM_THREADS, gemm_smem_bytes<32>(), stream>>>(A, mapA, B, mapB, M, N, K, out);
    } else {
        if (al) gemm_tn_kernel<128, EPI, true><<<grid, GEMM_THREADS, gemm_smem_bytes<128>(), stream>>>(A, mapA, B, mapB, M, N, K, out);
        else gemm_tn_kernel<128, EPI, false><<<grid, GEMM_THREADS, gemm_smem_bytes<128>(), stream>>>(A, mapA, B, mapB, M, N, K, out);
    }
    return cudaGetLastError();
}

template <int EPI>
inline cudaError_t gemm_tn_set_attributes()
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(gemm_tn_kernel<128, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<128>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemm_tn_kernel<128, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<128>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemm_tn_kernel<32, EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<32>())) != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tn_kernel<32, EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<32>());
}

}  // namespace fpt
