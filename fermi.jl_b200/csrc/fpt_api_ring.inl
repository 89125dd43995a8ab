// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: DF route with a ring of occupied slabs.

// ---- DF route without the full (ov|vv) block: a ring of occupied slabs (north star: "assembling (bd|ai) slices on the fly") ---------
// Pt holds o slabs of vp^2 Kp doubles (22.9 GB at C5; o = 100, v = 800 would need 467 GB).  Every slab is a GEMM away from the B
// factors (2 naux v^3 flops), so Pt need not be resident: the occupied range is cut into blocks of `ob`, the triplet list is walked
// block triple by block triple (I >= J >= K; i in I, j in J, k in K), and the ring holds just the 3 ob slabs of the current block
// triple -- slot group 0 for I, 1 for J (or I's when J = I), 2 for K (or J's when K = J).  Going to the next K assembles ob slabs;
// a block triple of distinct blocks carries ob^3 triplets of 12 v^3 (v + o) flops each, so the re-assembly costs
// naux / (6 ob^2 (v + o)) of the work: 4 % at C3 with ob = 4, 65 % with ob = 1 (three slabs in all).  Per block triple: one
// assembly launch group + one launch of the fused kernel over its explicit triplet list (Problem::trips) with the slot map of that
// block triple (Problem::pslot); E(T) accumulates on the device.  Several GPUs: everyone assembles the same slabs and takes its
// cost-weighted shard of every launch.
struct RingPhase { int I, J, K; i64 trip_off, ntrip; };

static int assemble_slabs(fpt_handle* h, Dev& d, const Problem& P, const int* pslot, int naux, const double* dBOV, const double* dBVV, int p0, int p1)
{
    if (p1 <= p0) return 0;
    const int o = P.o, v = P.v;
    prep_pt_hole<<<grid1d((i64)(p1 - p0) * o * v * v), 256, 0, d.stream>>>(P, d.Pt.d(), d.cur_T2, p0, p1 - p0, pslot);
    CK(cudaGetLastError());
    GemmOut out{};
    out.P = P;
    out.C = d.Pt.d();
    out.p0 = p0;
    out.pslot = pslot;
    const RowMap mA{p0, o, 1, v};   // m = y + v*pl  ->  BOV row (p0+pl) + o*y
    const RowMap mB{0, v, 1, v};    // n = d + v*x   ->  BVV row x + v*d
    CK(gemm_tn_launch<EPI_PT>(d.stream, dBOV, mA, dBVV, mB, (i64)(p1 - p0) * v, v * v, naux, out, d.n_sm));
    h->launches += 2;
    return 0;
}

static int triples_df_ring(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO, const double* BOV,
                           const double* BVV, const double* fo, const double* fv, int ob)
{
    const int nblk = (o + ob - 1) / ob, L = (int)h->devs.size();
    if (setup_problem(h, o, v, 3 * ob)) return 1;
    std::vector<const double*> dBOO, dBOV, dBVV;
    if (upload_t1_f(h, T1, fo, fv)) return 1;
    if (upload_t2(h, T2, {}, false)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOO; }, BOO, (size_t)naux * o * o, dBOO)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBOV; }, BOV, (size_t)naux * o * v, dBOV)) return 1;
    if (distribute(h, [](Dev& d) -> DevBuf& { return d.sBVV; }, BVV, (size_t)naux * v * v, dBVV)) return 1;
    // the phases: block triples in the order I, J <= I, K <= J; triplet lists in the reference's loop order inside each
    std::vector<RingPhase> phases;
    std::vector<int> trips, pslots;
    auto blk_lo = [&](int B) { return B * ob; };
    auto blk_hi = [&](int B) { return std::min(o, (B + 1) * ob); };
    for (int I = 0; I < nblk; I++)
        for (int J = 0; J <= I; J++)
            for (int K = 0; K <= J; K++) {
                RingPhase ph{I, J, K, (i64)trips.size() / 3, 0};
                for (int i = blk_lo(I); i < blk_hi(I); i++)
                    for (int j = blk_lo(J); j < blk_hi(J) && j <= i; j++)
                        for (int k = blk_lo(K); k < blk_hi(K) && k <= j; k++)
                            if (!(i == j && j == k)) { trips.push_back(i); trips.push_back(j); trips.push_back(k); ph.ntrip++; }
                if (ph.ntrip == 0) continue;
                // slot map of this block triple: I -> group 0, J -> group 1 unless J == I, K -> group 2 unless K == J (or I)
                std::vector<int> m((size_t)o, 0);
                const int gJ = (J == I) ? 0 : 1, gK = (K == J) ? gJ : 2;
                for (int p = blk_lo(I); p < blk_hi(I); p++) m[p] = 0 * ob + (p - blk_lo(I));
                for (int p = blk_lo(J); p < blk_hi(J); p++) m[p] = gJ * ob + (p - blk_lo(J));
                for (int p = blk_lo(K); p < blk_hi(K); p++) m[p] = gK * ob + (p - blk_lo(K));
                pslots.insert(pslots.end(), m.begin(), m.end());
                phases.push_back(ph);
            }
    i64 total = 0;
    for (const RingPhase& ph : phases) total += ph.ntrip;
    if (total != num_triplets(o)) return fail("internal: ring phases hold %lld triplets, expected %lld", (long long)total, (long long)num_triplets(o));
    for (int g = 0; g < L; g++) {
        Dev& d = *h->devs[g];
        CK(cudaSetDevice(d.dev));
        const Problem& P0 = d.prob;
        // Qt hole part and OV2 as on the materialised route                                       (DFERI.jl:88-112, 139-154)
        GemmOut out{};
        out.P = P0;
        out.C = d.Qt.d();
        CK(gemm_tn_launch<EPI_QT_HOLE>(d.stream, dBOO[g], rowmap_identity(), dBOV[g], rowmap_identity(), (i64)o * o, o * v, naux, out, d.n_sm));
        CK(cudaMemsetAsync(d.OV2.p, 0, (size_t)ov2_elems(P0) * sizeof(double), d.stream));
        out.C = d.OV2.d();
        CK(gemm_tn_launch<EPI_OV2>(d.stream, dBOV[g], rowmap_identity(), dBOV[g], rowmap_identity(), (i64)o * v, o * v, naux, out, d.n_sm));
        if (d.ringtab.ensure((trips.size() + pslots.size()) * sizeof(int))) return 1;
        int* dtr = (int*)d.ringtab.p;
        int* dps = dtr + trips.size();
        // pageable sources: the copies are staged by the driver before the call returns, the vectors may go out of scope
        CK(cudaMemcpyAsync(dtr, trips.data(), trips.size() * sizeof(int), cudaMemcpyHostToDevice, d.stream));
        CK(cudaMemcpyAsync(dps, pslots.data(), pslots.size() * sizeof(int), cudaMemcpyHostToDevice, d.stream));
        if (g == 0) { CK(cudaEventRecord(d.tl[1], d.copy)); CK(cudaEventRecord(d.tl[2], d.stream)); }
        CK(cudaEventRecord(d.ev0[0], d.stream));
        int curI = -1, curJ = -1, curK = -1;
        for (size_t t = 0; t < phases.size(); t++) {
            const RingPhase& ph = phases[t];
            Problem P = current_problem(h, d);
            const RingMap ring{dps + t * (size_t)o, dtr + 3 * ph.trip_off};
            P.tw_begin = 0;
            P.tw_count = ph.ntrip;
            P.nitems = P.nb * ph.ntrip;
            if (ph.I != curI) { if (assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.I), blk_hi(ph.I))) return 1; curI = ph.I; curJ = curK = -1; }
            if (ph.J != curJ) { if (ph.J != ph.I && assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.J), blk_hi(ph.J))) return 1; curJ = ph.J; curK = -1; }
            if (ph.K != curK) { if (ph.K != ph.J && assemble_slabs(h, d, P, ring.pslot, naux, dBOV[g], dBVV[g], blk_lo(ph.K), blk_hi(ph.K))) return 1; curK = ph.K; }
            i64 sb, se;
            shard_range(h, P, 0, P.nitems, d.grank, h->world, &sb, &se);
            if (compute_launch(h, d, P, sb, se, -1, t > 0, ring)) return 1;
        }
        CK(cudaEventRecord(d.ev1[0], d.stream));
        d.clean_o = -1;   // the ring's slabs are not the layout a later materialised upload expects to find clean
    }
    h->launches += 2 * (int)phases.size() - 2;   // compute_finish counts one kernel + one reduction per GPU itself
    h->nphase = 1;
    h->ring_call = true;
    h->last_profiled = false;
    return 0;
}

// Block size of the slab ring for this problem (0: materialise all o slabs).  fpt_set_df_ring: -1 never, 0 automatic -- ring
// with blocks of 4 when the full Pt would take more than 40 % of the device's memory --, n >= 1 ring with blocks of n.
static int df_ring_block(fpt_handle* h, int o, int v)
{
    if (h->dbg_flags || h->profiling || h->item_order != 1) return 0;
    if (h->df_ring > 0) return std::min(h->df_ring, o);
    if (h->df_ring < 0) return 0;
    const size_t total_b = h->devs[0]->total_mem;   // from the device properties: cudaMemGetInfo costs tens of milliseconds per call
    const double vp = padded_v(v), Kp = roundup(v + o, KGROUP);
    const double full = (double)o * vp * vp * Kp * sizeof(double);
    return (full > 0.4 * (double)total_b && o > 12) ? 4 : 0;
}

static int triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO, const double* BOV,
                      const double* BVV, const double* fo, const double* fv, double* Et, fpt_stats* st, bool async, const char* who)
{
    if (check_idle(h, who)) return 1;
    if (!T1 || !T2 || !BOO || !BOV || !BVV || !fo || !fv || (!async && !Et)) return fail("%s: NULL argument", who);
    if (naux < 1) return fail("%s: invalid naux=%d", who, naux);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, who, {T1, T2, BOO, BOV, BVV, fo, fv})) return 1;
    upload_begin(h);
    const int ob = df_ring_block(h, o, v);
    if (ob > 0) {
        if (triples_df_ring(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, ob)) return 1;
        h->last = fpt_stats{};
        h->last.h2d_bytes = h->h2d;
        if (compute_collect(h, h->nitems)) return 1;
        return finish_tail(h, async, t0, Et, st);
    }
    if (upload_df_impl(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, false)) return 1;
    return finish_call(h, async, t0, Et, st);
}

extern "C" int fpt_triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                                const double* OOOV, const double* OVOV, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    return triples_conv(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, Et, st, false, "fpt_triples_conv");
}
extern "C" int fpt_triples_conv_async(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV,
                                      const double* OOOV, const double* OVOV, const double* fo, const double* fv)
{
    return triples_conv(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, nullptr, nullptr, true, "fpt_triples_conv_async");
}
extern "C" int fpt_triples_df(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                              const double* BOV, const double* BVV, const double* fo, const double* fv, double* Et, fpt_stats* st)
{
    return triples_df(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, Et, st, false, "fpt_triples_df");
}
extern "C" int fpt_triples_df_async(fpt_handle* h, int o, int v, int naux, const double* T1, const double* T2, const double* BOO,
                                    const double* BOV, const double* BVV, const double* fo, const double* fv)
{
    return triples_df(h, o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, nullptr, nullptr, true, "fpt_triples_df_async");
}

extern "C" int fpt_wait(fpt_handle* h, double* Et, fpt_stats* st)
{
    if (!h || !Et) return fail("fpt_wait: NULL argument");
    if (!h->pending) return fail("fpt_wait: no asynchronous call is in flight");
    DeviceGuard guard;
    h->pending = false;
    const auto t0 = wall::now();
    if (compute_finish(h, Et, nullptr)) return 1;
    h->last.total_ms = h->last.upload_ms + ms_since(t0);   // host time inside the two calls
    if (st) *st = h->last;
    return 0;
}
