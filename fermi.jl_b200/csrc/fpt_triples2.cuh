// K1+K2, second generation: the fused RCCSD(T) kernel with the accumulate-into-W epilogue moved off the tensor-core warps.
//
// Same math, operand layouts, work list, producer and energy stage as fpt_triples.cuh (ijk.jl:108-136); what changes is who
// adds a finished GEMM's accumulators into the shared-memory W slots.  In the first kernel the 16 DMMA warps did it
// themselves between two k-loops (9 % of a C4 launch with the tensor pipe idle, plus the exposed latency of the next
// GEMM's first operand loads).  Here
//   * a consumer warp that finishes a k-loop *parks* its 16 accumulator doubles in tensor memory (one tcgen05.st of 32
//     columns, SASS STTM; TMEM is otherwise unused by an FP64 kernel) and goes straight on to the next k-loop;
//   * four epilogue warps -- one per TMEM lane quarter, so that lane l of epilogue warp q reads exactly what lane l of the
//     consumer warps q, q+4, q+8, q+12 wrote -- pull the parked values back (tcgen05.ld, LDTM) and add them into the W
//     slots, concurrently with the k-loops of the following GEMM.  The first contribution a slot receives is a plain
//     store (GemmDesc::dfirst), so the slots are never zeroed.
// Different GEMMs reach the same W element through different thread -> element maps, so the four epilogue warps order
// their "sets" (one (GEMM, column half) each) with a named barrier; the consumers only meet the epilogue again before the
// energy stage (`w_done`).
//
// CTA = 4 consumer warpgroups (96 registers) + 1 epilogue warpgroup (72) + 1 producer warpgroup (24) = 768 threads.
#pragma once
#include "fpt_triples.cuh"

namespace fpt {

constexpr int NEWARPS = 4;
constexpr int NTHREADS2 = NCTHREADS + 128 + 128;
constexpr int LAUNCH_REGS2 = 80, CONSUMER_REGS2 = 96, EPILOGUE_REGS2 = 72, PRODUCER_REGS2 = 24;
static_assert(NCTHREADS * CONSUMER_REGS2 + 128 * EPILOGUE_REGS2 + 128 * PRODUCER_REGS2 <= NTHREADS2 * LAUNCH_REGS2,
              "setmaxnreg budget exceeds the CTA pool");
constexpr int PARK_COLS = 32;                      // 32-bit TMEM columns per (consumer warp, buffer): 16 doubles
constexpr int PARK_BUFS = 2;
constexpr int TMEM_COLS = (NCWARPS / 4) * PARK_BUFS * PARK_COLS;   // 256 columns x 128 lanes
static_assert(TMEM_COLS == 256, "tcgen05.alloc takes a power of two");
constexpr int NPROF2 = 12;

struct SmemTail2 {
    Ctl ctl[2];
    unsigned long long full[QSTAGES], empty[QSTAGES], item_full[2], item_empty[2], ov_full, ov_empty;
    unsigned long long park_full[4][PARK_BUFS], park_empty[4][PARK_BUFS], w_done, w_free;
    double red[NCWARPS];
    uint32_t tmem_base, pad_[3];
};

constexpr size_t TRIPLES2_SMEM_BYTES = (size_t)(WSLOT_DOUBLES + QSTAGES * QSTAGE_DOUBLES) * sizeof(double) + sizeof(SmemTail2);
static_assert(TRIPLES2_SMEM_BYTES <= 232448, "more than the 227 KB a CTA may use");

// TMEM address of the parking area of consumer warp (quarter q, index cw in the quarter), buffer `buf`
__device__ __forceinline__ uint32_t park_addr(uint32_t tmem_base, int q, int cw, int buf)
{
    return tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((cw * PARK_BUFS + buf) * PARK_COLS);
}

// one GEMM of an item on a consumer warp: k-loop, next GEMM's row setup and first A loads, park the accumulators.
// MTW = number of row tiles this warp owns in this GEMM (0: none, it only keeps the ring and parking protocols in step).
template <int MTW, int NT, bool PROF>
__device__ __forceinline__ void gemm_body2(const Problem& P, const Ctl* ctl, int g, RowSet& rs, double4x (&a)[ABUF][MTW_MAX],
                                           const double* Qsm, SmemTail2* tail, int& stage, uint32_t& sphase, uint32_t& pcount,
                                           uint32_t tmem_base, int warp, int lane, long long* prof)
{
    const GemmDesc& gd = ctl->ent.gemm[g];
    double acc[MTW > 0 ? MTW : 1][NT][2];
    long long t0 = 0, t1 = 0;
    if (PROF) t0 = clock64();
    if constexpr (MTW > 0) gemm_kloop<MTW, NT, PROF>(P, gd, rs, a, acc, Qsm, tail, stage, sphase, lane, prof);
    else kloop_idle(P, tail, stage, sphase, lane);
    if (PROF) { t1 = clock64(); prof[2] += t1 - t0; }
    const bool dup_next = (g + 1 < ctl->ent.ngemm) && gemm_is_dup(ctl->item, g + 1);
    const int gnext = g + (dup_next ? 2 : 1);
    if (gnext < ctl->ent.ngemm) {
        const GemmDesc& gn = ctl->ent.gemm[gnext];
        rows_setup2<false>(P, RingMap{nullptr, nullptr}, gn, occ_pick(ctl->item, gn.p), warp, lane, rs);
        a_prologue2(P, rs, a);
    }
    if (PROF) { t0 = clock64(); prof[1] += t0 - t1; }
    const int q = warp & 3, cw = warp >> 2, buf = pcount & 1;
    mbar_wait((uint64_t*)&tail->park_empty[q][buf], ((pcount >> 1) & 1) ^ 1);   // the epilogue has drained this buffer
    if (PROF) { t1 = clock64(); prof[6] += t1 - t0; }
    if constexpr (MTW > 0) {
        tcgen05_fence_after();
        park_acc<MTW, NT>(acc, park_addr(tmem_base, q, cw, buf));
        tcgen05_fence_before();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive((uint64_t*)&tail->park_full[q][buf]);
    pcount++;
    if (PROF) prof[3] += clock64() - t1;
}

// ---------------------------------------------------------------------------------------------------
// epilogue warp of TMEM lane quarter q: adds the parked accumulators of consumer warps q, q+4, q+8, q+12 into the W slots
// ---------------------------------------------------------------------------------------------------
template <bool PROF>
__device__ __forceinline__ void epilogue_loop(const Problem& P, double* Wsm, SmemTail2* tail, int q, int lane, uint32_t tmem_base,
                                              long long* prof_out)
{
    const int kk = lane & 3, r = lane >> 2;
    uint32_t pcount = 0;
    long long ep[6] = {0, 0, 0, 0, 0, 0};
    long long tp = 0, tstart = 0;
    if (PROF) tstart = clock64();
    for (uint32_t n = 0;; n++) {
        const int slot = n & 1;
        const Ctl* ctl = &tail->ctl[slot];
        if (PROF) tp = clock64();
        mbar_wait((uint64_t*)&tail->item_full[slot], (n >> 1) & 1);
        if (ctl->cur_item < 0) break;
        if (PROF) { const long long t = clock64(); ep[0] += t - tp; tp = t; }
        if (n > 0) mbar_wait((uint64_t*)&tail->w_free, (n - 1) & 1);   // the consumers have finished the previous energy stage
        if (PROF) ep[1] += clock64() - tp;
        const int ngemm = ctl->ent.ngemm;
        for (int g = 0; g < ngemm; g++) {
            if (gemm_is_dup(ctl->item, g)) continue;
            const int ndup = ((g + 1 < ngemm) && gemm_is_dup(ctl->item, g + 1)) ? 1 : 0;
            const int buf = pcount & 1;
            if (PROF) tp = clock64();
            mbar_wait((uint64_t*)&tail->park_full[q][buf], (pcount >> 1) & 1);
            if (PROF) ep[2] += clock64() - tp;
            tcgen05_fence_after();
            for (int rep = 0; rep <= ndup; rep++) {   // a twin GEMM (i == j or j == k) adds the same accumulators at its own places
                const GemmDesc& gd = ctl->ent.gemm[g + rep];
                const int TX = gd.TX, xinv = gd.xinv, rt_total = gd.rt_total;
                const int nt = gd.TZ >> 2;
#pragma unroll 1
                for (int e = 0; e < 2; e++) {
                    if (PROF) tp = clock64();
                    named_bar_sync(2, NEWARPS * 32);   // every epilogue warp has finished the previous set
                    if (PROF) { const long long t = clock64(); ep[3] += t - tp; tp = t; }
                    const int dbase = gd.dbase[e], sel = gd.dsel[e], Tb = gd.dTb[e], Tc = gd.dTc[e];
                    const bool first = gd.dfirst[e] != 0;
                    // which of (x, y, z) supplies (la, lb, lc); z = 4*ct + kk is the coordinate that runs over the column tiles
                    const int ra = sel & 3, rb = (sel >> 2) & 3, rc = (sel >> 4) & 3;
                    const int zs = (rc == 2) ? 0 : ((rb == 2) ? 4 * Tc : 4 * Tb * Tc);
                    const int d1 = (rc == 2) ? 4 : 1;
                    const int mask = (Tc == 16) ? 15 : 3;
#pragma unroll 1
                    for (int cw = 0; cw < NCWARPS / 4; cw++) {
                        int rt0, nv;
                        warp_rows(rt_total, q + 4 * cw, rt0, nv);
                        if (nv == 0) continue;   // warp-uniform
                        uint32_t v[16];
                        if (!(P.dbg_flags & 32)) tmem_ld16(park_addr(tmem_base, q, cw, buf) + (uint32_t)(e * 16), v);   // (mt 0..1) x (ct 0..3) doubles
                        else { for (int t = 0; t < 16; t++) v[t] = 0u; }
                        int off[2][4];
#pragma unroll
                        for (int mt = 0; mt < 2; mt++) {
                            const int m = (rt0 + mt) * 8 + r;
                            const int yl = (m * xinv) >> 16;
                            const int xl = m - yl * TX;
                            const int la = pick3(ra, xl, yl, kk), lb = pick3(rb, xl, yl, kk), lc = pick3(rc, xl, yl, kk);
                            const int lin0 = dbase + (la * Tb + lb) * Tc;
                            const int w0 = lc ^ ((swz_a(la) ^ swz_b(lb)) & mask);
#pragma unroll
                            for (int ct = 0; ct < 4; ct++) off[mt][ct] = lin0 + ct * zs + (w0 ^ (ct * d1));
                        }
                        tmem_wait_ld();
                        if (!(P.dbg_flags & 1)) {
                            double d[2][4];
#pragma unroll
                            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                                for (int ct = 0; ct < 4; ct++) d[mt][ct] = __hiloint2double((int)v[(mt * 4 + ct) * 2 + 1], (int)v[(mt * 4 + ct) * 2]);
                            if (!first) {
                                double o[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
                                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                                    for (int ct = 0; ct < 4; ct++)
                                        if (mt < nv && ct < nt) o[mt][ct] = Wsm[off[mt][ct]];
#pragma unroll
                                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                                    for (int ct = 0; ct < 4; ct++) {
                                        if (P.dbg_flags & 16)   // timing experiment: no FP64 add (integer xor keeps the dependency)
                                            d[mt][ct] = __longlong_as_double(__double_as_longlong(d[mt][ct]) ^ __double_as_longlong(o[mt][ct]));
                                        else
                                            d[mt][ct] += o[mt][ct];
                                    }
                            }
#pragma unroll
                            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                                for (int ct = 0; ct < 4; ct++)
                                    if (mt < nv && ct < nt) Wsm[off[mt][ct]] = d[mt][ct];
                        }
                    }
                    if (PROF) ep[4] += clock64() - tp;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive((uint64_t*)&tail->park_empty[q][buf]);
            pcount++;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive((uint64_t*)&tail->w_done);             // this warp's share of every set of the item is in the slots
            mbar_arrive((uint64_t*)&tail->item_empty[slot]);
        }
    }
    if (PROF && q == 0 && lane == 0) {   // the epilogue's view replaces the second consumer observer
        ep[5] = clock64() - tstart;
        for (int t = 0; t < 6; t++) prof_out[blockIdx.x * NPROF + 12 + t] = ep[t];
        for (int t = 6; t < 12; t++) prof_out[blockIdx.x * NPROF + 12 + t] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// The kernel.  PROF (warp 0 and warp 12): {wait-for-item, next-GEMM setup, k-loops, park, energy, total, park-buffer wait,
// wait for the epilogue before the energy stage, Q-ring wait inside the k-loops, OV2 tile wait, -, barrier after energy}
// ---------------------------------------------------------------------------------------------------
template <bool PROF>
__global__ void __launch_bounds__(NTHREADS2, 1)
triples_kernel2(Problem P, i64 item_begin, i64 item_end, unsigned long long* counter, double* partials, long long* prof_out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Wsm = reinterpret_cast<double*>(smem_raw);
    double* Qsm = Wsm + WSLOT_DOUBLES;
    SmemTail2* tail = reinterpret_cast<SmemTail2*>(Qsm + QSTAGES * QSTAGE_DOUBLES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < QSTAGES; s++) { mbar_init((uint64_t*)&tail->full[s], 1); mbar_init((uint64_t*)&tail->empty[s], NCWARPS); }
        for (int s = 0; s < 2; s++) { mbar_init((uint64_t*)&tail->item_full[s], 1); mbar_init((uint64_t*)&tail->item_empty[s], NCWARPS + NEWARPS); }
        mbar_init((uint64_t*)&tail->ov_full, 1);
        mbar_init((uint64_t*)&tail->ov_empty, NCWARPS);
        for (int qq = 0; qq < 4; qq++)
            for (int b = 0; b < PARK_BUFS; b++) {
                mbar_init((uint64_t*)&tail->park_full[qq][b], NCWARPS / 4);
                mbar_init((uint64_t*)&tail->park_empty[qq][b], 1);
            }
        mbar_init((uint64_t*)&tail->w_done, NEWARPS);
        mbar_init((uint64_t*)&tail->w_free, NCWARPS);
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == NCWARPS) tmem_alloc<TMEM_COLS>(&tail->tmem_base);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp >= NCWARPS + NEWARPS) {   // producer warpgroup
        setmaxnreg_dec<PRODUCER_REGS2>();
        if (warp == NCWARPS + NEWARPS && lane == 0) producer_loop<false, false>(P, RingMap{nullptr, nullptr}, item_begin, item_end, counter, Qsm, tail);
        return;
    }
    if (warp >= NCWARPS) {             // epilogue warpgroup
        setmaxnreg_dec<EPILOGUE_REGS2>();
        epilogue_loop<PROF>(P, Wsm, tail, warp - NCWARPS, lane, tmem_base, prof_out);
        named_bar_sync(2, NEWARPS * 32);   // every epilogue warp is done with TMEM (the consumers' last park was read above)
        if (warp == NCWARPS) tmem_dealloc<TMEM_COLS>(tmem_base);
        return;
    }
    setmaxnreg_inc<CONSUMER_REGS2>();

    // ------------------------------- consumers -------------------------------
    long long prof[NPROF2] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_start = 0;
    if (PROF) t_start = clock64();
    double esum = 0.0;
    int stage = 0;
    uint32_t sphase = 0, pcount = 0;
    double4x a[ABUF][MTW_MAX];

    for (uint32_t n = 0;; n++) {
        const int slot = n & 1;
        const Ctl* ctl = &tail->ctl[slot];
        long long t0 = 0, t1 = 0;
        if (PROF) t0 = clock64();
        mbar_wait((uint64_t*)&tail->item_full[slot], (n >> 1) & 1);
        if (ctl->cur_item < 0) break;
        if (PROF) { t1 = clock64(); prof[0] += t1 - t0; }

        const int ngemm = ctl->ent.ngemm;
        RowSet rs;
        rows_setup2<false>(P, RingMap{nullptr, nullptr}, ctl->ent.gemm[0], occ_pick(ctl->item, ctl->ent.gemm[0].p), warp, lane, rs);
        a_prologue2(P, rs, a);
        if (PROF) { t0 = clock64(); prof[1] += t0 - t1; }

        for (int g = 0; g < ngemm; g++) {
            if (gemm_is_dup(ctl->item, g)) continue;   // its twin's accumulators are added a second time by the epilogue
            const GemmDesc& gd = ctl->ent.gemm[g];
            const int nt = gd.TZ >> 2;
            // rs (set up during the previous GEMM) holds this warp's share of row tiles: 0, 1 or 2
            switch (__shfl_sync(0xffffffffu, rs.nvalid, 0)) {   // broadcast: warp-uniform for the compiler (see fpt_triples.cuh)
            case 0: FPT_DISPATCH_NT(0, nt, (gemm_body2<MTW, NT, PROF>(P, ctl, g, rs, a, Qsm, tail, stage, sphase, pcount, tmem_base, warp, lane, prof))) break;
            case 1: FPT_DISPATCH_NT(1, nt, (gemm_body2<MTW, NT, PROF>(P, ctl, g, rs, a, Qsm, tail, stage, sphase, pcount, tmem_base, warp, lane, prof))) break;
            default: FPT_DISPATCH_NT(2, nt, (gemm_body2<MTW, NT, PROF>(P, ctl, g, rs, a, Qsm, tail, stage, sphase, pcount, tmem_base, warp, lane, prof))) break;
            }
        }
        if (PROF) t0 = clock64();
        mbar_wait((uint64_t*)&tail->w_done, n & 1);           // every contribution of this item is in the W slots
        if (PROF) { t1 = clock64(); prof[7] += t1 - t0; t0 = t1; }
        {
            const BlockDesc& bd = ctl->ent.bd;
            const int TC = bd.ts[2];
            const int half = tid >> 8, tt = tid & 255;   // two threads per (b,c) column, 8 values of a each
            mbar_wait((uint64_t*)&tail->ov_full, n & 1);      // the 12 a-row OV2 tiles are in the ring area
            if (PROF) prof[9] += clock64() - t0;
            if (!(P.dbg_flags & 2)) {
                if (bd.slot_elems == 4096)
                    esum += block_column_energy_t<true>(P, bd, ctl->item.i, ctl->item.j, ctl->item.k, Wsm, Qsm, tt >> 4, tt & 15,
                                                        half * 8, half * 8 + 8);
                else if (tt < bd.ts[1] * TC)
                    esum += block_column_energy_t<false>(P, bd, ctl->item.i, ctl->item.j, ctl->item.k, Wsm, Qsm, tt / TC, tt % TC,
                                                         half * 8, half * 8 + 8);
            }
        }
        if (PROF) t1 = clock64();
        consumer_bar();       // the staged tiles and ctl[slot] may be reused; the W slots may be overwritten
        if (PROF) prof[11] += clock64() - t1;
        if (lane == 0) {
            mbar_arrive((uint64_t*)&tail->ov_empty);
            mbar_arrive((uint64_t*)&tail->item_empty[slot]);
            mbar_arrive((uint64_t*)&tail->w_free);
        }
        if (PROF) prof[4] += clock64() - t0;
    }

    // CTA reduction (warp shuffle, then one thread sums the warp partials in fixed order)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, off);
    if (lane == 0) tail->red[warp] = esum;
    consumer_bar();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < NCWARPS; w++) s += tail->red[w];
        partials[blockIdx.x] = s;
    }
    if (PROF && lane == 0 && warp == 0) {
        prof[5] = clock64() - t_start;
        for (int t = 0; t < NPROF2; t++) prof_out[blockIdx.x * NPROF + t] = prof[t];
    }
}

}  // namespace fpt
