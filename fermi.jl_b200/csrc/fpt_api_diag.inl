// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: calibration and diagnostics.

// ---- calibration and diagnostics ---------------------------------------------------------------------------------------------
extern "C" int fpt_fp64_peak(fpt_handle* h, int variant, double ms_target, double* tflops)
{
    if (!h || !tflops) return fail("fpt_fp64_peak: NULL argument");
    if (check_idle(h, "fpt_fp64_peak")) return 1;
    DeviceGuard guard;
    Dev& d = *h->devs[0];
    CK(cudaSetDevice(d.dev));
    if (d.out.ensure(OUT_DOUBLES * sizeof(double))) return 1;
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
    if (d.out.ensure(OUT_DOUBLES * sizeof(double))) return 1;
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
