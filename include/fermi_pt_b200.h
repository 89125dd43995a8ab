/*
 * fermi_pt_b200.h -- C ABI of libfermi_pt_b200.so, the B200-native RCCSD(T) perturbative-triples engine.
 *
 * The reference (Fermi.jl, pure Julia) has no FFI for this path; its boundary is the Julia method
 *     RCCSDpT(ccsd::RCCSD, moints::IntegralHelper{T,E,O}, Alg::ijk)        src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150
 * selected by dispatch on an `RpTAlgorithm` singleton                     .../PerturbativeTriples.jl:1-11,51-63
 * The entry points below are exactly what a Julia `ccall` (or Python ctypes) binding of that method needs:
 * plain pointers and sizes, no torch / CUDA types.  The Julia glue that binds them is in
 * fermi.jl_b200/julia/FermiB200.jl and is described in INTEGRATION.md.
 *
 * Array arguments are the reference's own arrays, Float64, Julia column-major (first index fastest):
 *     T1  [i,a]      (o,v)        ccsd.T1            RCCSD.jl:65
 *     T2  [i,j,a,b]  (o,o,v,v)    ccsd.T2            RCCSD.jl:66
 *     OVVV[i,a,b,c]  (o,v,v,v)    moints["OVVV"]     ijk.jl:28   (= (ia|bc))
 *     OOOV[i,j,k,a]  (o,o,o,v)    moints["OOOV"]     ijk.jl:31   (= (ij|ka))
 *     OVOV[i,a,j,b]  (o,v,o,v)    moints["OVOV"]     ijk.jl:32   (= (ia|jb))
 *     fo  [i]        (o)          moints["Fii"]      ijk.jl:36
 *     fv  [a]        (v)          moints["Faa"]      ijk.jl:37
 *     BOO [Q,i,j]    (naux,o,o)   moints["BOO"]      DFERI.jl:15-29
 *     BOV [Q,i,a]    (naux,o,v)   moints["BOV"]      DFERI.jl:31-51
 *     BVV [Q,a,b]    (naux,v,v)   moints["BVV"]      DFERI.jl:53-69
 * Pointers may be host memory (pageable or pinned) or device memory; they are read-only, caller-owned, and only need to stay
 * valid for the duration of the call.
 *   - pageable host memory (what a Julia ccall or numpy hands over) is copied by a pool of host threads into a ring of pinned
 *     bounce buffers and sent from there, so it moves at the PCIe rate; with several GPUs every GPU pulls only its own 1/N of
 *     each array over its own PCIe link and the parts are exchanged with one ncclAllGather over NVLink.  The arrays carry the
 *     index symmetries OVVV[i,a,b,c] = OVVV[i,a,c,b], T2[i,j,a,b] = T2[j,i,b,a], OVOV[i,a,j,b] = OVOV[j,b,i,a] (the reference's
 *     own algorithms agree with each other only then); the library reads just the half those leave free (b <= c, a <= b) and
 *     writes the mirror images on the GPU, which halves the bytes the host has to move.  Each array is spot-checked first and
 *     read in full if it does not look symmetric; fpt_set_symmetric_inputs(h, 0) reads everything in full;
 *   - device memory must live on the handle's (first) GPU and is consumed in place on the library's own streams: the library
 *     issues one cudaDeviceSynchronize() on that GPU before the first read, so work queued on ANY stream of the caller
 *     (PyTorch's or CUDA.jl's current stream) that produced the arrays is complete -- no caller-side synchronisation needed.
 * Every entry point restores the caller's current CUDA device before it returns.
 *
 * Every function returns 0 on success and a nonzero status on failure; fpt_last_error() then describes it
 * (the Julia glue rethrows it as FermiException, Options.jl:196-199).  There is no CPU fallback: if no
 * usable GPU is present fpt_create fails.
 */
#ifndef FERMI_PT_B200_H
#define FERMI_PT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpt_handle fpt_handle;

typedef struct fpt_stats {
    double upload_ms;      /* host->device copies + layout prep (K4) or DF assembly (K3) of the last upload */
    double kernel_ms;      /* fused triples kernel of the last compute, CUDA events on the launch stream */
    double total_ms;       /* host wall clock of the last fpt_triples_* call */
    double flops;          /* algorithmic flops of the last compute: 12 v^3 (v+o) per non-zero-weight triplet share */
    double h2d_bytes;      /* bytes copied host->device by the last upload */
    long long n_items;     /* work items (triplet x block) processed by the last compute */
    long long n_triplets;  /* non-zero-weight triplets (i>=j>=k, not i=j=k) of the problem */
    int n_launches;        /* kernels launched by the last upload + compute */
    int n_sm;              /* SMs of the device */
} fpt_stats;

/* ngpu = 1: devices[0] = CUDA device ordinal (devices == NULL -> current device).
 * ngpu > 1 (single-process multi-GPU, what a Julia caller uses): devices[0..ngpu) form one NCCL clique (libnccl.so.2 is
 * loaded at run time).  Every upload is sharded: GPU g copies part g of every large array from the caller's memory over its
 * own PCIe link, one in-place ncclAllGather per array (per 64 MB chunk of OVVV) completes it on every GPU over NVLink, and
 * every GPU runs the layout prep itself; every compute splits its item range into ngpu contiguous cost-weighted shards and
 * ends with one scalar ncclAllReduce of E(T). */
int fpt_create(int ngpu, const int* devices, fpt_handle** out);
int fpt_destroy(fpt_handle* h);

/* One process per GPU (torchrun / MPI style launchers): rank 0 calls fpt_nccl_unique_id, the launcher's control plane hands
 * the 128 bytes to the other ranks, and every rank calls fpt_create_rank(device, rank, world, id, &h)  (ncclCommInitRank).
 * On such a handle fpt_upload_*, fpt_triples_* and fpt_compute are COLLECTIVE calls: every rank makes the same call with
 * the same arrays (host arrays: every rank reads only its own 1/world share of them) and item range; the data plane
 * -- sharded H2D, all-gather, all-reduce -- is the library's own, and every rank receives the full E(T). */
int fpt_nccl_unique_id(void* id128);
int fpt_create_rank(int device, int rank, int world, const void* id128, fpt_handle** out);
/* host threads that copy pageable memory into the pinned ring (default: the process's CPU affinity count, divided by `world`
 * on rank handles, at most 16; environment override FERMI_PT_B200_THREADS) */
int fpt_set_host_threads(fpt_handle* h, int n);
/* on (default): pageable host inputs cross PCIe as their symmetry-unique halves (see the top of this file); 0: always in full */
int fpt_set_symmetric_inputs(fpt_handle* h, int on);

/* replaces RCCSDpT(ccsd, moints, ::ijk) for conventional integrals (ijk.jl:20-150): Et = E(T) */
int fpt_triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                     const double* OOOV, const double* OVOV, const double* fo, const double* fv, double* Et,
                     fpt_stats* stats);

/* same, density-fitted: the (ia|bd), (ij|ka), (ia|jb) blocks that DFERI.jl:88-180 would materialise on the
 * host are assembled on the device from the B factors */
int fpt_triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                   const double* BOV, const double* BVV, const double* fo, const double* fv, double* Et,
                   fpt_stats* stats);

/* Handles over several GPUs split the work list into shards of equal ESTIMATED cost.  With on = 1 (default) repeated calls of one
 * shape refine the split from the shards' measured kernel times, which travel with the scalar all-reduce of the next call: after every
 * second call the boundaries move to where equal times lie (measured at C4 on 8 GPUs: shard times within +-2 % before, see DESIGN.md
 * for after).  0: always the model's split.  Rank handles: set it identically on every rank (it sizes the all-reduce). */
int fpt_set_adaptive_shards(fpt_handle* h, int on);

/* E(T) is a sum of per-CTA partial sums added in a fixed order; by default the CTAs pull work items off a global counter, so which
 * CTA sums which items -- and with it the last bits of E(T), ~1e-16 relative -- varies from run to run.  on = 1: items are dealt
 * statically (CTA b takes items b, b + grid, ...): the same call on the same GPU type returns the same bits every time, at the cost of
 * the dynamic balance (measured at C4: see DESIGN.md). */
int fpt_set_deterministic(fpt_handle* h, int on);

/* Slab ring of the density-fitted route: instead of all o slabs (p.|..) -- o vp^2 Kp doubles, 22.9 GB at o = 40, v = 400 -- only the
 * 3 * block slabs of the current block triple of occupied indices are resident and are re-assembled from the B factors on the fly
 * (cost: naux / (6 block^2 (v + o)) of the (T) work).  block = -1: never; 0 (default): automatic, blocks of 4 when the full set would
 * take more than 40 % of the device memory; n >= 1: always, with blocks of n (n = 1: three slabs in all).  Applies to fpt_triples_df
 * and fpt_triples_df_async; fpt_upload_df always materialises.  fpt_device_bytes: device memory the handle holds on its first GPU. */
int fpt_set_df_ring(fpt_handle* h, int block);
int fpt_device_bytes(fpt_handle* h, double* bytes);

/* Single-precision callers (`@set precision single` makes every array of the reference Float32, IntegralHelper.jl:58-68): the same
 * arrays as above in 4-byte form, host memory.  They cross PCIe as they are (half the bytes), are widened on the GPU and evaluated in
 * FP64: Et is the exact (T) energy of the rounded inputs. */
int fpt_triples_conv_f32(fpt_handle* h, int o, int v, const float* T1, const float* T2, const float* OVVV, const float* OOOV,
                         const float* OVOV, const float* fo, const float* fv, double* Et, fpt_stats* stats);
int fpt_triples_df_f32(fpt_handle* h, int o, int v, int naux, const float* T1, const float* T2, const float* BOO, const float* BOV,
                       const float* BVV, const float* fo, const float* fv, double* Et, fpt_stats* stats);

/* Asynchronous forms (gradient_findif makes 6 N_atoms (T) calls, FiniteDifferences.jl:48-74): the call returns
 * as soon as the caller's arrays have been consumed -- they may be freed or overwritten, the next CCSD can start on the CPU --
 * while the GPU is still computing; fpt_wait blocks for E(T).  One call may be in flight per handle. */
int fpt_triples_conv_async(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                           const double* OOOV, const double* OVOV, const double* fo, const double* fv);
int fpt_triples_df_async(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                         const double* BOV, const double* BVV, const double* fo, const double* fv);
int fpt_wait(fpt_handle* h, double* Et, fpt_stats* stats);

/* Same, from the AO-basis integrals: replaces the reference's dense AO -> MO contractions for the three blocks the (T)
 * path reads (Chonky.jl:28-48 OOOV, :72-92 OVOV, :94-114 OVVV) with four quarter transformations on the GPU, so the o*v^3
 * block never exists on the host.  AOERI = aoints["ERI"], nbf^4 column-major (mu|nu rho sigma chemist order as stored by
 * the reference); Co = C[:, (1+drop_occ):ndocc] (nbf x o) and Cv = C[:, (ndocc+1):(nbf-drop_vir)] (nbf x v), the slices
 * Chonky.jl:38-41 takes. */
int fpt_triples_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                   const double* Co, const double* Cv, const double* fo, const double* fv, double* Et, fpt_stats* stats);
int fpt_upload_ao(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, const double* AOERI,
                  const double* Co, const double* Cv, const double* fo, const double* fv);

/* Same, from the reference's default AO container for conventional integrals: the sparse list of symmetry-unique
 * (mu nu|rho sigma) values (FermiSparse, Arrays.jl:9-12, built in AtomicIntegrals.jl:48-52) that Sparse.jl:78-151 (OOOV),
 * :236-313 (OVOV), :316-393 (OVVV) scatter and contract on the CPU.  vals = aoints["ERI"].data (nint doubles), idx =
 * aoints["ERI"].indexes: nint zero-based 4-tuples stored contiguously as index_bytes-wide signed integers (2 for
 * Vector{NTuple{4,Int16}}, 4 for Int32).  The list is expanded on the GPU into the dense tensor (nbf^4 doubles must fit the
 * device), then the fpt_triples_ao path runs. */
int fpt_triples_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                          const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                          const double* fo, const double* fv, double* Et, fpt_stats* stats);
int fpt_upload_ao_sparse(fpt_handle* h, int nbf, int o, int v, const double* T1, const double* T2, long long nint,
                         const void* idx, int index_bytes, const double* vals, const double* Co, const double* Cv,
                         const double* fo, const double* fv);

/* ---- beside the (T) path ----
 * DF-CCSD particle-particle ladder, the hot spot of a DF-CCSD iteration: replaces
 *     cc_update_T2_v4_term!(newT2, T1, T2, moints::IntegralHelper{T,<:AbstractDFERI}, ::RCCSDa)      RCCSDHelper.jl:204-220
 * newT2[i,j,a,b] += sum_cd (T2[i,j,c,d] + T1[i,c] T1[j,d]) sum_Q BVV[Q,c,a] BVV[Q,d,b].  newT2 (o,o,v,v) is host memory, read and
 * updated in place; T1, T2, BVV as above.  The (vv|vv) block is assembled slab by slab on the GPU and never stored. */
int fpt_ccsd_ladder_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BVV,
                       double* newT2, fpt_stats* stats);
/* MP2 correlation energy: replaces RMP2_energy for density-fitted (RMP2a.jl:91-143) and conventional (RMP2a.jl:146-169) integrals,
 * E = sum_iajb (ia|jb) [2 (ia|jb) - (ib|ja)] / (fo[i] + fo[j] - fv[a] - fv[b]). */
int fpt_mp2_df(fpt_handle* h, int o, int v, int naux, const double* BOV, const double* fo, const double* fv, double* Emp2,
               fpt_stats* stats);
int fpt_mp2_conv(fpt_handle* h, int o, int v, const double* OVOV, const double* fo, const double* fv, double* Emp2, fpt_stats* stats);

/* Staged form of the calls above (kernel-only timing, partial evaluations):
 * upload = copy + layout prep, operands stay resident on the GPU(s); compute = fused kernel over the item range
 * [item_begin, item_end) of the static work list (item_end < 0: to the end), returning that range's share of E(T).  A handle
 * over several GPUs (fpt_create with ngpu > 1, or fpt_create_rank) splits the range into one cost-weighted shard per GPU and
 * all-reduces the scalar. */
int fpt_upload_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                    const double* OOOV, const double* OVOV, const double* fo, const double* fv);
int fpt_upload_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                  const double* BOV, const double* BVV, const double* fo, const double* fv);
int fpt_num_items(fpt_handle* h, long long* n_items);
int fpt_compute(fpt_handle* h, long long item_begin, long long item_end, double* Et_partial, fpt_stats* stats);

/* The static work list.  An item is (triplet i >= j >= k, tile triple of the virtual range); the list holds every item of
 * the triplets in the current window.
 *  - fpt_set_triplet_window: restrict the list to positions [t_begin, t_end) of the reference's own flattened triplet list
 *    (the order its loops walk, ijk.jl:49,63,83: i, then j <= i, then k <= j fastest; zero-weight i = j = k entries count as
 *    positions but do no work).  t_end < 0 = to the end.  A new upload resets the window to all triplets.
 *  - fpt_set_item_order: 1 (default) block-major -- all triplets of one tile triple before the next, which keeps that tile
 *    triple's operand panels L2-resident; 0 triplet-major.
 *  - fpt_shard_items: part `rank` of `world` of the list as a contiguous item range of equal estimated cost -- what each
 *    GPU computes (one process per GPU passes it to fpt_compute; a multi-GPU handle does the same split internally). */
int fpt_set_triplet_window(fpt_handle* h, long long t_begin, long long t_end);
int fpt_set_item_order(fpt_handle* h, int order);
int fpt_shard_items(fpt_handle* h, int rank, int world, long long* item_begin, long long* item_end);

/* FP64 pipe calibration for the roofline denominator: variant 0 = DMMA.8x8x4 stream, 1 = DFMA stream.
 * Returns sustained TFLOP/s over `ms_target` milliseconds of back-to-back launches. */
int fpt_fp64_peak(fpt_handle* h, int variant, double ms_target, double* tflops);
/* Stand-alone timing of the DF-assembly / quarter-transform GEMM (fpt_gemm.cuh) on zero operands: C(M x N) = A(M x K).B(N x K)^T,
 * `reps` launches, sustained TFLOP/s. */
int fpt_gemm_bench(fpt_handle* h, long long M, int N, int K, int reps, double* tflops);
/* Milliseconds since the start of the last upload on the first GPU's clock: out8 = {host time spent copying pageable memory
 * into the pinned ring, last H2D done, operands ready (gathers + prep done), kernel begin, kernel end, result sent, 0, 0}. */
int fpt_last_timeline(fpt_handle* h, double* out8);

/* Diagnostics (not part of the drop-in path): with profiling on, fpt_compute runs the instrumented kernel variant and
 * fpt_last_profile returns its phase breakdown in SM cycles summed over CTAs (warp 0's view) --
 * out24 = {wait-for-item, zero, k-loops, RMW epilogues (+token wait, next-GEMM prologue), energy stage, total, token wait,
 *          pure RMW, wait on the Q ring inside the k-loops, wait for the staged OV2 tiles, barrier before / after the energy stage}
 * for the first warp of consumer group 0 (entries 0-11) and of group 3 (entries 12-23); and a DMMA issue study. */
int fpt_set_profiling(fpt_handle* h, int on);
int fpt_set_kernel_variant(fpt_handle* h, int variant);   /* 1 (default): the DMMA warps update the W slots themselves; 2: experimental epilogue-warp kernel (TMEM parking), only in builds with -DFPT_WITH_VARIANT2 */
int fpt_set_debug_flags(fpt_handle* h, int flags);   /* 1: skip RMW, 2: skip energy stage -- timing studies only, E(T) is wrong */
int fpt_last_profile(fpt_handle* h, double* out24);
int fpt_dmma_sweep(fpt_handle* h, int ilp, int warps_per_sm, double* tflops);

const char* fpt_last_error(void);
const char* fpt_version(void);

#ifdef __cplusplus
}
#endif
#endif
