// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: beside the (T) path: Float32 callers, the DF-CCSD particle-particle ladder, the MP2 energy.

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
        d.out.ensure(OUT_DOUBLES * sizeof(double)))
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
