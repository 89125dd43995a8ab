// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: fused-kernel launches, shards, collection of E(T); the one-call forms (split call, asynchronous calls).

// ---- compute ---------------------------------------------------------------------------------------------------------------
static Problem current_problem(const fpt_handle* h, const Dev& d)
{
    Problem P = d.prob;
    P.order = h->item_order;
    P.dbg_flags = h->dbg_flags | (h->deterministic ? 512 : 0);
    P.tw_begin = h->tw_begin;
    P.tw_count = h->tw_count;
    P.nitems = h->nitems;
    return P;
}

// Static split of the item range [b, e) into `world` contiguous parts of equal estimated cost (shard_items in fpt_layout.h,
// shared with the CPU emulator so that the gloo tests exercise the very same split)
static void shard_range(const fpt_handle* h, const Problem& P, i64 b, i64 e, int rank, int world, i64* sb, i64* se, const double* frac = nullptr)
{
    shard_items(P, h->block_cost.data(), b, e, rank, world, sb, se, frac);
}

// Boundary fractions of phase `phase` for the split of [b, e) (see ShardCal): uniform until two calls with the same boundaries
// have delivered every GPU's kernel time, then moved (damped) to where the measured time density puts equal times.
static const double* shard_fractions(fpt_handle* h, const Problem& P, i64 b, i64 e, int phase)
{
    if (h->world < 2 || !h->adaptive || phase < 0 || phase >= MAX_PHASES || h->profiling || h->dbg_flags) return nullptr;
    ShardCal& c = h->cal[phase];
    const int W = h->world;
    const bool same = c.o == P.o && c.v == P.v && c.world == W && c.order == P.order && c.tw_begin == P.tw_begin &&
                      c.tw_count == P.tw_count && c.b == b && c.e == e && (int)c.frac.size() == W + 1;
    if (!same) {
        c = ShardCal{};
        c.o = P.o; c.v = P.v; c.world = W; c.order = P.order; c.tw_begin = P.tw_begin; c.tw_count = P.tw_count; c.b = b; c.e = e;
        c.frac.resize(W + 1);
        for (int r = 0; r <= W; r++) c.frac[r] = (double)r / W;
        c.gen = ++h->gen_counter;   // unique per handle: a time measured with other boundaries (or another problem) never matches
    } else if (c.pending) {
        std::vector<double> nf(W + 1, 0.0);
        if (rebalance_fractions(W, c.frac.data(), c.ms.data(), 0.8, nf.data())) { c.frac = nf; c.gen = ++h->gen_counter; }
        c.pending = false;
    }
    return c.frac.data();
}

// launch the fused kernel + reduction for [item_begin, item_end) of the work list described by P on one GPU (asynchronous; the result
// is stored in d.out, or added to it for the second phase of a split call, which also has its own pair of timing events)
static int compute_launch(fpt_handle* h, Dev& d, const Problem& P, i64 item_begin, i64 item_end, int phase, int accumulate = -1,
                          RingMap ring = RingMap{nullptr, nullptr})
{
    // phase >= 0: the launch is timed with event pair `phase`; accumulate (default: phase > 0): add to d.out instead of storing
    if (accumulate < 0) accumulate = phase > 0;
    CK(cudaSetDevice(d.dev));
    const i64 n = item_end - item_begin;
    int grid = d.n_sm;
    if ((i64)grid > n) grid = (int)(n > 0 ? n : 1);
    CK(cudaMemsetAsync(d.counter.p, 0, sizeof(unsigned long long), d.stream));
    if (phase >= 0) CK(cudaEventRecord(d.ev0[phase], d.stream));
    unsigned long long* ctr = (unsigned long long*)d.counter.p;
#ifdef FPT_WITH_VARIANT2
    if (h->kernel_variant == 2) {
        if (h->profiling) triples_kernel2<true><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p);
        else triples_kernel2<false><<<grid, NTHREADS2, TRIPLES2_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p);
    } else
#endif
    if (ring.trips)   // slab ring of the DF route: explicit triplet list, Pt through the slot map
        triples_kernel<false, true><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    else if (h->profiling)
        triples_kernel<true><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    else
        triples_kernel<false><<<grid, NTHREADS, TRIPLES_SMEM_BYTES, d.stream>>>(P, item_begin, item_end, ctr, d.partials.d(), (long long*)d.prof.p, ring);
    d.last_grid = grid;
    CK(cudaGetLastError());
    if (phase >= 0) CK(cudaEventRecord(d.ev1[phase], d.stream));
    {   // this GPU's kernel time of the previous call (if it was measured with the boundaries still in use) rides along, see ShardCal
        const bool cal = phase >= 0 && h->world > 1 && h->adaptive && !ring.trips;
        const int W = h->world;
        const double prev = (cal && d.last_gen[phase] == h->cal[phase].gen) ? d.last_ms[phase] : 0.0;
        reduce_partials<<<1, 32, 0, d.stream>>>(d.partials.d(), grid, d.out.d(), accumulate, cal ? MAX_PHASES * W : 0,
                                                cal ? phase * W + d.grank : -1, prev);
    }
    CK(cudaGetLastError());
    d.shard_b = item_begin;
    d.shard_e = item_end;
    return 0;
}

// Enqueue the kernels for items [item_begin, item_end) of the work list over the triplet window [tw_begin, tw_begin + tw_count):
// every GPU of the communicator takes its static, cost-weighted shard.
static int compute_launch_all(fpt_handle* h, i64 tw_begin, i64 tw_count, i64 item_begin, i64 item_end, int phase)
{
    const double* frac = nullptr;
    for (Dev* dp : h->devs) {
        Problem P = current_problem(h, *dp);
        P.tw_begin = tw_begin;
        P.tw_count = tw_count;
        P.nitems = P.nb * tw_count;
        if (dp == h->devs[0]) frac = shard_fractions(h, P, item_begin, item_end, phase);   // the same for every GPU of the handle
        i64 sb, se;
        shard_range(h, P, item_begin, item_end, dp->grank, h->world, &sb, &se, frac);
        if (compute_launch(h, *dp, P, sb, se, phase)) return 1;
    }
    h->last_profiled = h->profiling;
    return 0;
}

// E(T) is one scalar all-reduce; the 8-byte result is sent to the host.  Nothing here waits for the GPU.
static int compute_collect(fpt_handle* h, i64 n_items)
{
    if (h->world > 1) {
        NCK(nccl_api().GroupStart());
        const size_t count = h->adaptive ? 1 + (size_t)MAX_PHASES * h->world : 1;   // E(T) + the time slots (ShardCal)
        for (Dev* dp : h->devs) NCK(nccl_api().AllReduce(dp->out.p, dp->out.p, count, ncclDouble, ncclSum, dp->comm, dp->stream));
        NCK(nccl_api().GroupEnd());
    }
    Dev& d0 = *h->devs[0];
    CK(cudaSetDevice(d0.dev));
    CK(cudaMemcpyAsync(h->res_pinned, d0.out.p, (h->world > 1 && h->adaptive ? OUT_DOUBLES : 1) * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    CK(cudaEventRecord(d0.tl[5], d0.stream));
    h->pend_items = n_items;
    return 0;
}

static int compute_enqueue(fpt_handle* h, i64 item_begin, i64 item_end)
{
    if (item_end < 0 || item_end > h->nitems) item_end = h->nitems;
    if (item_begin < 0) item_begin = 0;
    if (item_begin > item_end) item_begin = item_end;
    h->nphase = 1;
    if (compute_launch_all(h, h->tw_begin, h->tw_count, item_begin, item_end, 0)) return 1;
    return compute_collect(h, item_end - item_begin);
}

static int compute_finish(fpt_handle* h, double* Et, fpt_stats* st)
{
    float ms_max = 0.f;
    for (Dev* dp : h->devs) {
        CK(cudaSetDevice(dp->dev));
        CK(cudaStreamSynchronize(dp->stream));
        float sum = 0.f;
        for (int t = 0; t < h->nphase; t++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, dp->ev0[t], dp->ev1[t]));
            sum += ms;
            dp->last_ms[t] = ms;                         // travels with the next call's all-reduce (ShardCal)
            dp->last_gen[t] = h->cal[t].gen;
        }
        if (sum > ms_max) ms_max = sum;
    }
    if (Et) *Et = *h->res_pinned;
    if (h->world > 1 && h->adaptive && !h->ring_call) {
        // the reduced vector holds, per phase, every GPU's kernel time of the PREVIOUS call where that call used the boundaries
        // still in use (else 0): once all are there, the next launch of the phase moves its boundaries
        const int W = h->world;
        for (int t = 0; t < h->nphase; t++) {
            ShardCal& c = h->cal[t];
            if ((int)c.frac.size() != W + 1 || c.pending) continue;
            bool all = true;
            for (int r = 0; r < W; r++) all = all && h->res_pinned[1 + t * W + r] > 0.0;
            if (all) {
                c.ms.assign(h->res_pinned + 1 + t * W, h->res_pinned + 1 + (t + 1) * W);
                c.pending = true;
            }
        }
    }
    // algorithmic flops of the triplets in the window, scaled by the share of the window's items that were computed
    const double ntrip = (double)h->tw_count;
    const int v = h->v, o = h->o;
    h->last.kernel_ms = ms_max;
    h->last.n_items = h->pend_items;
    h->last.n_triplets = (long long)ntrip;
    h->last.flops = 12.0 * v * (double)v * v * (v + o) * ntrip * (h->nitems ? (double)h->pend_items / (double)h->nitems : 0.0);
    h->last.n_launches = h->launches + 2 * h->nphase * (int)h->devs.size();
    h->last.n_sm = h->devs[0]->n_sm;
    // timeline of the first GPU, milliseconds since the upload began (entries stay 0 when the compute followed an older upload)
    Dev& d0 = *h->devs[0];
    CK(cudaSetDevice(d0.dev));
    for (double& t : h->timeline) t = 0.0;
    h->timeline[0] = h->stage_host_ms;
    cudaEvent_t marks[5] = {d0.tl[1], d0.tl[2], d0.ev0[0], d0.ev1[h->nphase - 1], d0.tl[5]};
    for (int t = 0; t < 5; t++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d0.tl[0], marks[t]) == cudaSuccess) h->timeline[1 + t] = ms;
        else cudaGetLastError();
    }
    if (st) *st = h->last;
    return 0;
}

extern "C" int fpt_compute(fpt_handle* h, long long item_begin, long long item_end, double* Et, fpt_stats* st)
{
    if (!h || !Et) return fail("fpt_compute: NULL argument");
    if (check_idle(h, "fpt_compute")) return 1;
    if (!h->loaded) return fail("fpt_compute: no problem uploaded");
    DeviceGuard guard;
    if (compute_enqueue(h, item_begin, item_end)) return 1;
    return compute_finish(h, Et, st);
}

// Milliseconds since the start of the last upload, on the first GPU's clock: out8 = {host time spent copying pageable memory into
// the pinned ring (wall, overlaps the DMAs), last H2D done, operands ready (gathers + prep done), kernel begin, kernel end,
// result on its way to the host, 0, 0}.
extern "C" int fpt_last_timeline(fpt_handle* h, double* out8)
{
    if (!h || !out8) return fail("fpt_last_timeline: NULL argument");
    for (int t = 0; t < 8; t++) out8[t] = h->timeline[t];
    return 0;
}

// ---- one-call forms ----------------------------------------------------------------------------------------------------------
// upload and compute are enqueued back to back (no host synchronisation in between); `async` returns as soon as the caller's
// arrays have been consumed, fpt_wait collects the result.
static int finish_tail(fpt_handle* h, bool async, wall::time_point t0, double* Et, fpt_stats* st)
{
    if (async) {
        for (Dev* dp : h->devs) {   // inputs in pinned memory are read by the DMA engines directly: wait for those reads
            CK(cudaSetDevice(dp->dev));
            CK(cudaStreamSynchronize(dp->copy));
        }
        h->last.upload_ms = ms_since(t0);
        h->pending = true;
        return 0;
    }
    if (compute_finish(h, Et, nullptr)) return 1;
    h->last.total_ms = ms_since(t0);
    h->last.upload_ms = h->timeline[2];
    if (st) *st = h->last;
    return 0;
}
static int finish_call(fpt_handle* h, bool async, wall::time_point t0, double* Et, fpt_stats* st)
{
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    if (compute_enqueue(h, 0, -1)) return 1;
    return finish_tail(h, async, t0, Et, st);
}

// Phases of a one-call conventional evaluation.  The triplets with i < pb only read the operands of the occupied indices p < pb, and
// they are the first num_triplets(pb) entries of the reference's triplet list (ijk.jl:49,63,83 loops i slowest).  So the call is cut
// at occupied boundaries 0 = pb[0] < pb[1] < ... < pb[n] = o:  upload everything but OVVV, then OVVV[p < pb[1]]; launch the kernel over
// that window; while it runs, the host threads and the DMA engines bring OVVV[pb[1] <= p < pb[2]]; and so on.  The kernel starts after
// 1/n of OVVV has arrived, and the rest of the host-bound staging time -- the part of an 8-GPU call that does not shrink with the
// number of GPUs -- disappears behind the kernels: phase t holds (pb[t+1]^3 - pb[t]^3) / o^3 of the work, which covers the staging
// of slice t+1 as long as the whole kernel takes longer than the whole upload.
// Conditions: host-resident OVVV small enough that the later slices (packed copies on every GPU) are cheap to hold, enough occupied
// orbitals, a full default work list.  Boundaries are multiples of 4 (32-byte rows for the streaming copies).
// FERMI_PT_B200_SPLIT = number of phases wanted (default 2; 0 or 1: no split; more phases start the kernel earlier but re-read
// more of the host array's cache lines -- a slice is 8 (o / nph) bytes of every 8 o-byte row; measured at C4: no gain beyond 2).
static int split_points(fpt_handle* h, int o, int v, const double* T2, const double* OVVV, const double* OVOV, int* pb)
{
    int want = 2;
    if (const char* e = getenv("FERMI_PT_B200_SPLIT")) want = atoi(e);
    if (want > MAX_PHASES) want = MAX_PHASES;
    pb[0] = 0;
    pb[1] = o;
    if (want < 2 || o < 8 || classify(OVVV) == PK_DEVICE) return 1;
    if ((double)o * v * v * v * sizeof(double) > 4e9) return 1;
    if (h->dbg_flags || h->profiling || h->item_order != 1) return 1;
    int n = 0;
    for (int t = 1; t < want; t++) {
        const int b = (int)(((i64)o * t / want + 2) & ~3);
        if (b > pb[n] && b < o) pb[++n] = b;
    }
    pb[++n] = o;
    return n;
}

static int triples_conv(fpt_handle* h, int o, int v, const double* T1, const double* T2, const double* OVVV, const double* OOOV,
                        const double* OVOV, const double* fo, const double* fv, double* Et, fpt_stats* st, bool async, const char* who)
{
    if (check_idle(h, who)) return 1;
    if (!T1 || !T2 || !OVVV || !OOOV || !OVOV || !fo || !fv || (!async && !Et)) return fail("%s: NULL argument", who);
    DeviceGuard guard;
    const auto t0 = wall::now();
    if (admit_device_inputs(h, who, {T1, T2, OVVV, OOOV, OVOV, fo, fv})) return 1;
    upload_begin(h);
    int pb[MAX_PHASES + 1];
    const int nph = split_points(h, o, v, T2, OVVV, OVOV, pb);
    if (nph <= 1) {
        if (upload_conv_impl(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, false)) return 1;
        return finish_call(h, async, t0, Et, st);
    }
    h->nphase = nph;
    auto launch_phase = [&](int t) -> int {
        if (t == nph - 1 && upload_end(h, false)) return 1;
        const i64 nb = h->devs[0]->prob.nb;
        const i64 u0 = num_triplets(pb[t]), u1 = num_triplets(pb[t + 1]);
        return compute_launch_all(h, u0, u1 - u0, 0, nb * (u1 - u0), t);
    };
    if (upload_conv_slices(h, o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, pb, nph, launch_phase)) return 1;
    const i64 nb = h->devs[0]->prob.nb;
    h->last = fpt_stats{};
    h->last.h2d_bytes = h->h2d;
    if (compute_collect(h, nb * num_triplets(o))) return 1;
    return finish_tail(h, async, t0, Et, st);
}
