// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: problem set-up and the uploads: layout prep, symmetry-unique halves, conventional slices, DF assembly.

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
        if (d.out.ensure(OUT_DOUBLES * sizeof(double))) return 1;
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
    h->ring_call = false;
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
        const View vw = view_t2_half(o, v);
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
        const View vw = view_ovov_half(o, v);   // for every (b, j): the prefix a <= b of the (i, a) plane
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
            const View vw = view_ovvv_chunk(o, v, p0, np, c0, cn, half);
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
