// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: AO -> MO routes (dense AO tensor, sparse AO list).

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
