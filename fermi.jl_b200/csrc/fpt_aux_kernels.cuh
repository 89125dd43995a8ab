// Auxiliary sm_100a kernels of the B200 RCCSD(T) path:
//   K4  prep_*            layout prep (ijk.jl:24-32 replaced by one pass into the Pt/Qt/OV2/T1d layouts)
//   K6  expand_sparse_eri_kernel   sparse AO list -> dense AO tensor
//   (K3 / K5, the DF assembly and the AO -> MO quarter transforms, are the GEMM of fpt_gemm.cuh)
//   reduce_partials       fixed-order final sum (ijk.jl:145)
//   peak_* / dmma_ilp_*   FP64 pipe calibration (roofline denominator)
#pragma once
#include <cuda_runtime.h>
#include "fpt_layout.h"
#include "fpt_ptx.cuh"

namespace fpt {

// out[0] = (accumulate ? out[0] : 0) + sum of the partials.  out[1 .. nslots] carry one number per (phase, GPU) through the scalar
// all-reduce of a multi-GPU handle -- this GPU's kernel time of the previous call, for the adaptive shard balance: a store
// launch (accumulate = 0) clears them, every launch writes its own slot (slot < 0: none).
__global__ void reduce_partials(const double* partials, int n, double* out, int accumulate, int nslots = 0, int slot = -1, double value = 0.0)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = accumulate ? out[0] : 0.0;
        for (int t = 0; t < n; t++) s += partials[t];
        out[0] = s;
        if (!accumulate)
            for (int t = 0; t < nslots; t++) out[1 + t] = 0.0;
        if (slot >= 0) out[1 + slot] = value;
    }
}

// ---------------------------------------------------------------------------------------------------
// K4: layout prep.  Sources are the reference's column-major arrays (first index fastest).
// ---------------------------------------------------------------------------------------------------
// Pt[p][y][x][d] = OVVV[p,y,x,d] for d in [d0, d0+dn) and p in [p0, p0+np); `src` holds that sub-block compactly,
// src[pl + np*(y + v*(x + v*dd))] (np = o, p0 = 0: it points at OVVV[:,:,:,d0] itself).
__global__ void prep_pt_particle(Problem P, double* Pt, const double* __restrict__ src, int d0, int dn, int p0, int np)
{
    __shared__ double tile[32][33];
    const int o = np, v = P.v;
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
            const int p = p0 + (int)(py % o), y = (int)(py / o);
            Pt[pt_row(P, p, y, x) + d0 + dd] = tile[tx][ty + kk];
        }
    }
}

// Pt[slot(p)][y][x][v+l] = -T2[p,l,y,x] for p in [p0, p0 + np)   (pslot: slab ring of the DF route, else nullptr)
__global__ void prep_pt_hole(Problem P, double* Pt, const double* __restrict__ T2, int p0, int np, const int* __restrict__ pslot)
{
    const int o = P.o, v = P.v;
    const i64 n = (i64)np * v * v * o;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int l = (int)(idx % o);
        i64 t = idx / o;
        const int x = (int)(t % v); t /= v;
        const int y = (int)(t % v);
        const int p = p0 + (int)(t / v);
        Pt[pt_row_slot(P, pslot, p, y, x) + v + l] = -T2[p + (i64)o * (l + (i64)o * (y + (i64)v * x))];
    }
}

// Qt[(q,r)][g][z][kk8]: kappa<v: T2[r,q,z,kappa]; v<=kappa<v+o: OOOV[kappa-v,q,r,z]; else 0.
// OOOV == nullptr (density-fitted route): the hole part is left zero here and written by the DF assembly GEMM.
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
            else if (kappa < v + o && OOOV) val = OOOV[(kappa - v) + (i64)o * (q + (i64)o * (r + (i64)o * z))];
        }
        Qt[idx] = val;
    }
}

// OV2[(q,r)][Y][Z][16][16] = OVOV[q,y,r,z] (zero padded to 16-multiples)
__global__ void prep_ov2(Problem P, double* OV2, const double* __restrict__ OVOV)
{
    const int o = P.o, v = P.v, nt = P.nt;
    const i64 n = ov2_elems(P);
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int zl = (int)(idx & 15), yl = (int)((idx >> 4) & 15);
        i64 t = idx >> 8;
        const int Z = (int)(t % nt); t /= nt;
        const int Y = (int)(t % nt); t /= nt;
        const int r = (int)(t % o);
        const int q = (int)(t / o);
        const int y = tile_start(Y, P.vp) + yl, z = tile_start(Z, P.vp) + zl;
        const bool ok = yl < tile_size(Y, P.vp) && zl < tile_size(Z, P.vp) && y < v && z < v;
        OV2[idx] = ok ? OVOV[q + (i64)o * (y + (i64)v * (r + (i64)o * z))] : 0.0;
    }
}

// ---- symmetry-unique halves (host inputs cross PCIe as the half the index symmetry leaves free, fpt_api.cu) -----------------------
FPT_HD i64 tri(i64 c) { return c * (c + 1) / 2; }

// OVVV[i,a,b,c] = OVVV[i,a,c,b]: the packed source holds, for c in [c0, c0+cn), the prefix b <= c of the slice p in [p0, p0+np):
//     src[np v (tri(c) - tri(c0)) + pl + np (y + v b)] = OVVV[p0+pl, y, b, c].
// mirror = 0: Pt[p][y][x=b][kappa=c] (one block row per b, tiled transpose over (p y, c));
// mirror = 1: Pt[p][y][x=c][kappa=b] for b < c (one block row per c, tiled transpose over (p y, b)).
__global__ void prep_pt_particle_tri(Problem P, double* Pt, const double* __restrict__ src, int c0, int cn, int p0, int np, int mirror)
{
    __shared__ double tile[32][33];
    const int v = P.v;
    const i64 npv = (i64)np * v;
    const i64 py0 = (i64)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (!mirror) {
        const int b = blockIdx.z, cc0 = blockIdx.y * 32;
        if (c0 + cc0 + 31 < b) return;
        for (int kk = 0; kk < 32; kk += 8) {
            const int cc = cc0 + ty + kk, c = c0 + cc;
            const i64 py = py0 + tx;
            if (cc < cn && c >= b && py < npv) tile[ty + kk][tx] = src[npv * (tri(c) - tri(c0) + b) + py];
        }
        __syncthreads();
        for (int kk = 0; kk < 32; kk += 8) {
            const i64 py = py0 + ty + kk;
            const int cc = cc0 + tx, c = c0 + cc;
            if (cc < cn && c >= b && py < npv) Pt[pt_row(P, p0 + (int)(py % np), (int)(py / np), b) + c] = tile[tx][ty + kk];
        }
    } else {
        const int c = c0 + blockIdx.z, b0 = blockIdx.y * 32;
        if (b0 >= c) return;
        for (int kk = 0; kk < 32; kk += 8) {
            const int b = b0 + ty + kk;
            const i64 py = py0 + tx;
            if (b < c && py < npv) tile[ty + kk][tx] = src[npv * (tri(c) - tri(c0) + b) + py];
        }
        __syncthreads();
        for (int kk = 0; kk < 32; kk += 8) {
            const i64 py = py0 + ty + kk;
            const int b = b0 + tx;
            if (b < c && py < npv) Pt[pt_row(P, p0 + (int)(py % np), (int)(py / np), c) + b] = tile[tx][ty + kk];
        }
    }
}

// T2[i,j,a,b] = T2[j,i,b,a]: src holds a <= b, src[o^2 (tri(b) + a) + i + o j] = T2[i,j,a,b]; writes the full array
__global__ void expand_t2_tri(int o, int v, double* __restrict__ T2, const double* __restrict__ src)
{
    const i64 o2 = (i64)o * o, n = o2 * v * v;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int i = (int)(idx % o), j = (int)((idx / o) % o);
        const i64 t = idx / o2;
        const int a = (int)(t % v), b = (int)(t / v);
        T2[idx] = a <= b ? src[o2 * (tri(b) + a) + i + (i64)o * j] : src[o2 * (tri(a) + b) + j + (i64)o * i];
    }
}

// OVOV[i,a,j,b] = OVOV[j,b,i,a]: src holds a <= b, src[o^2 tri(b) + j o (b+1) + i + o a] = OVOV[i,a,j,b]; writes the full array
__global__ void expand_ovov_tri(int o, int v, double* __restrict__ OVOV, const double* __restrict__ src)
{
    const i64 o2 = (i64)o * o, n = o2 * v * v;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int i = (int)(idx % o);
        i64 t = idx / o;
        const int a = (int)(t % v); t /= v;
        const int j = (int)(t % o), b = (int)(t / o);
        OVOV[idx] = a <= b ? src[o2 * tri(b) + (i64)j * o * (b + 1) + i + (i64)o * a] : src[o2 * tri(a) + (i64)i * o * (a + 1) + j + (i64)o * b];
    }
}

// Float32 inputs (`@set precision single`, IntegralHelper.jl:58-68): the arrays cross PCIe as they are and are widened on the device
__global__ void widen_f32_kernel(double* __restrict__ dst, const float* __restrict__ src, i64 n)
{
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) dst[idx] = (double)src[idx];
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
__global__ void __launch_bounds__(1024) dmma_ilp_kernel(double* out, int iters, double seed)
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

namespace fpt {

// ---------------------------------------------------------------------------------------------------
// K6: sparse AO integral list -> dense AO tensor.  The reference's default AO container is a list of the symmetry-unique
// non-negligible (mu nu|rho sigma) with zero-based index 4-tuples (FermiSparse, Arrays.jl:9-12; filled by
// AtomicIntegrals.jl:48-52); its MO builders scatter every entry to its up-to-8 permutational images while contracting
// the first index (Sparse.jl:78-151, 236-313, 316-393).  Here the images are written into a zero-initialised dense tensor
// (plain stores: images of one entry that coincide carry the same value), which then feeds the K5 chain.
// ---------------------------------------------------------------------------------------------------
// An index outside [0, nbf) (e.g. a one-based list) is not stored: the entry is skipped and *bad is raised, which the host turns
// into an error return.
template <typename Ti>
__global__ void expand_sparse_eri_kernel(double* __restrict__ AO, const Ti* __restrict__ idx, const double* __restrict__ vals,
                                         i64 nint, int nbf, int* bad)
{
    const i64 n1 = nbf, n2 = n1 * nbf, n3 = n2 * nbf;
    for (i64 z = (i64)blockIdx.x * blockDim.x + threadIdx.x; z < nint; z += (i64)gridDim.x * blockDim.x) {
        const i64 m = idx[4 * z], n = idx[4 * z + 1], r = idx[4 * z + 2], s = idx[4 * z + 3];
        if (m < 0 || n < 0 || r < 0 || s < 0 || m >= nbf || n >= nbf || r >= nbf || s >= nbf) { *bad = 1; continue; }
        const double V = vals[z];
        AO[m + n1 * n + n2 * r + n3 * s] = V;
        AO[n + n1 * m + n2 * r + n3 * s] = V;
        AO[m + n1 * n + n2 * s + n3 * r] = V;
        AO[n + n1 * m + n2 * s + n3 * r] = V;
        AO[r + n1 * s + n2 * m + n3 * n] = V;
        AO[s + n1 * r + n2 * m + n3 * n] = V;
        AO[r + n1 * s + n2 * n + n3 * m] = V;
        AO[s + n1 * r + n2 * n + n3 * m] = V;
    }
}

}  // namespace fpt

namespace fpt {

// ---------------------------------------------------------------------------------------------------
// Section 8(f) rows beside the (T) path: DF-CCSD particle-particle ladder and MP2 energy
// ---------------------------------------------------------------------------------------------------
// tauT[(c,d) + v^2 (i + o j)] = T2[i,j,c,d] + T1[i,c] T1[j,d]      (RCCSDHelper.jl:208, stored with the contracted pair (c,d) fastest)
__global__ void ladder_tau_kernel(int o, int v, double* __restrict__ tauT, const double* __restrict__ T1, const double* __restrict__ T2)
{
    __shared__ double tile[32][33];
    const i64 o2 = (i64)o * o, v2 = (i64)v * v;
    const i64 ij0 = (i64)blockIdx.x * 32, cd0 = (i64)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int kk = 0; kk < 32; kk += 8) {
        const i64 cd = cd0 + ty + kk, ij = ij0 + tx;
        if (cd < v2 && ij < o2) {
            const int i = (int)(ij % o), j = (int)(ij / o), c = (int)(cd % v), d = (int)(cd / v);
            tile[ty + kk][tx] = T2[ij + o2 * cd] + T1[i + (i64)o * c] * T1[j + (i64)o * d];
        }
    }
    __syncthreads();
    for (int kk = 0; kk < 32; kk += 8) {
        const i64 ij = ij0 + ty + kk, cd = cd0 + tx;
        if (cd < v2 && ij < o2) tauT[cd + v2 * ij] = tile[tx][ty + kk];
    }
}

// E(MP2) = sum_{i,a,j,b} (ia|jb) [2 (ia|jb) - (ib|ja)] / (e_i + e_j - e_a - e_b)        (RMP2a.jl:155-166; the DF form RMP2a.jl:118-133
// is the same sum over i <= j with weight 2).  One block sums a fixed stretch of the index space, a second launch adds the block
// partials in order: the result does not depend on scheduling.
__global__ void __launch_bounds__(256) mp2_energy_kernel(int o, int v, const double* __restrict__ OVOV, const double* __restrict__ fo,
                                                         const double* __restrict__ fv, double* __restrict__ partials)
{
    __shared__ double red[8];
    const i64 ov = (i64)o * v, n = ov * ov;
    double e = 0.0;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (i64)gridDim.x * blockDim.x) {
        const int i = (int)(idx % o);
        i64 t = idx / o;
        const int a = (int)(t % v); t /= v;
        const int j = (int)(t % o), b = (int)(t / o);
        const double x = OVOV[idx], y = OVOV[i + (i64)o * (b + (i64)v * (j + (i64)o * a))];
        e += x * (2.0 * x - y) / (fo[i] + fo[j] - fv[a] - fv[b]);
    }
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w];
        partials[blockIdx.x] = s;
    }
}

}  // namespace fpt
