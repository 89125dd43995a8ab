// K1+K2: the fused, persistent, warp-specialised RCCSD(T) kernel for sm_100a.
//
//   W build   ijk.jl:108-115 (== ijk2.jl:110-153)  FP64 tensor-core DMMA.8x8x4, P rows streamed global->registers,
//                                                  Q rows staged by TMA bulk copies (cp.async.bulk + mbarrier)
//   6-fold permutation + V + denominators + energy  ijk.jl:116-136, on the W slots held in shared memory;
//                                                  W and V never go to HBM
//   accumulation                                   ijk.jl:133,145: per-thread FP64 partial, warp-shuffle, per-CTA partial
//
// CTA = 4 consumer warpgroups (16 warps, 112 registers/thread after setmaxnreg) + 1 producer warpgroup (24 registers),
// 640 threads, 1 CTA / SM, grid = #SMs.
//   producer (warp 16, lane 0): pulls items off the global counter, decodes them into a double-buffered control
//       block (item / block / GEMM descriptors) and streams the Q operand chunks of all GEMMs of the item through a
//       3-stage shared-memory ring with TMA bulk copies that complete on `full` mbarriers; stages are recycled
//       through `empty` mbarriers, so it runs ahead of the consumers across GEMM and item boundaries.
//   consumers: for each P-stationary GEMM  D[TX*TY x 2*TZ] = P_p . [Q_qr | Q_rq]  every warp owns <= 2 row tiles
//       (8 rows) x all column tiles; A fragments come straight from global memory (each P row is used by exactly one
//       warp, kappa-contiguous, 32 B per lane, 256-bit loads prefetched one 16-kappa group ahead), B fragments from the ring.  The
//       accumulators are then added into the W slots (swizzled, see fpt_layout.h), and after the last GEMM the
//       energy of the block's a>=b>=c points is evaluated from the slots.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "fpt_layout.h"
#include "fpt_ptx.cuh"

namespace fpt {

constexpr int NCWARPS = 16;                        // consumer warps = 4 warpgroups
constexpr int NCTHREADS = NCWARPS * 32;            // 512
constexpr int NTHREADS = NCTHREADS + 128;          // + producer warpgroup (only its first lane works)
// setmaxnreg only redistributes the CTA's launch-time allocation: 640 threads x 96 registers (launch bound) = 61440,
// so 512*R_consumer + 128*R_producer must not exceed that (otherwise the TRY_ALLOC spins forever).  LAUNCH_REGS is what ptxas
// assigns under the launch bound below; build.py checks the built library (cuobjdump -res-usage) and refuses anything else.
constexpr int LAUNCH_REGS = 96, CONSUMER_REGS = 112, PRODUCER_REGS = 24;
static_assert(NCTHREADS * CONSUMER_REGS + 128 * PRODUCER_REGS <= NTHREADS * LAUNCH_REGS, "setmaxnreg budget exceeds the CTA pool");
constexpr int QSTAGES = 3;
constexpr int QBLK = (TMAX + 1) * KGROUP + 2;      // doubles per (group, s) block: TZ rows of 16 kappa, skewed by 16 B mod 128 B
                                                   // so that the s=0 / s=1 halves of a quarter-warp hit disjoint banks
static_assert(OV_STAGE_TILES == QSTAGES * OV_TILES_PER_STAGE && OV_TILES_PER_STAGE * 256 <= CHUNK_GROUPS * 2 * QBLK &&
                  OV_STAGE_STRIDE == CHUNK_GROUPS * 2 * QBLK,
              "the energy stage's OV2 tiles are staged in the Q ring area, four per ring stage");
constexpr int QSTAGE_DOUBLES = CHUNK_GROUPS * 2 * QBLK;                // 1096 doubles = 8768 B
constexpr int WSLOT_DOUBLES = MAX_SLOTS * TMAX * TMAX * TMAX;          // 24576 doubles = 192 KB
constexpr int MTW_MAX = 2;
constexpr int APREF = 1;                           // A-fragment prefetch distance in kappa groups (of 16)
constexpr int ABUF = 2;                            // group gg lives in buffer gg % ABUF = gl % ABUF
static_assert(CHUNK_GROUPS % ABUF == 0 && APREF < ABUF, "A-fragment ring is indexed by the group's position in its chunk");

struct Ctl {
    i64 cur_item;          // < 0: no more work
    ItemDesc item;
    BlockTabEntry ent;     // bd, ngemm, gemm[] -- copied from the device block table by TMA; gemm[].p/q/r are positions in (i,j,k)
};
static_assert(sizeof(Ctl) % 16 == 0 && offsetof(Ctl, ent) % 16 == 0, "Ctl::ent is a TMA bulk-copy destination");
constexpr int NPROF = 24;   // 12 phase counters for each of two observer warps

struct SmemTail {
    Ctl ctl[2];
    unsigned long long full[QSTAGES], empty[QSTAGES], item_full[2], item_empty[2], rmw_done[2], ov_full, ov_empty;
    double red[NCWARPS];
};

constexpr size_t TRIPLES_SMEM_BYTES = (size_t)(WSLOT_DOUBLES + QSTAGES * QSTAGE_DOUBLES) * sizeof(double) + sizeof(SmemTail);

__device__ __forceinline__ void consumer_bar() { named_bar_sync(1, NCTHREADS); }

// RMW phases of *different* GEMMs may touch the same W elements (through different thread->element maps), RMW phases of
// the same GEMM never do (except the diag_xz aliasing handled inside gemm_rmw).  So all warps add their accumulators
// concurrently, and split mbarriers order the GEMMs: the n-th slot update of the launch ("event" n, counted by `gcount`
// identically in every warp) arrives on rmw_done[n & 1]; its phase number there is n >> 1.
//   * accumulate path: before event n a warp waits for event n-1 (every warp's RMW of the previous GEMM is done) -- a whole
//     k-loop after that RMW, so the wait practically never blocks;
//   * load-accumulate-store path (blocks of three distinct tiles, triplets without twin GEMMs; BlockTabEntry::forder): the
//     accumulators *start* from the slot contents (LDS, or zero for a slot's first contribution) and are *stored* back after
//     the k-loop.  No FP64 add, no load->add->store chain between two k-loops; since consecutive GEMMs of that order write
//     different slots, a warp only waits for event n-2 before it loads.

// ---------------------------------------------------------------------------------------------------
// producer: one thread
// ---------------------------------------------------------------------------------------------------
// Fetch the next item, decode it arithmetically and pull its block's descriptor table entry into ctl->ent with one TMA
// bulk copy that completes on item_full (the scalar fields are ordinary stores released by the same barrier).
template <bool RING>
__device__ __forceinline__ void producer_decode(const Problem& P, const RingMap& ring, i64 item_begin, i64 item_end, unsigned long long* counter,
                                                Ctl* ctl, uint64_t* item_full, uint32_t& nfetch)
{
    // dynamic: the next item off the global counter (which CTA sums which items then varies from run to run, and E(T) with it in the
    // last bits); deterministic (fpt_set_deterministic, dbg_flags bit 512): CTA b takes items b, b + grid, b + 2 grid, ... -- in the
    // block-major list neighbouring items cost the same, so the static deal is balanced up to the block boundaries
    const i64 it = (P.dbg_flags & 512) ? item_begin + blockIdx.x + (i64)(nfetch++) * gridDim.x : item_begin + (i64)atomicAdd(counter, 1ULL);
    if (it >= item_end) {
        ctl->cur_item = -1;
        mbar_arrive(item_full);
        return;
    }
    i64 block;
    item_decode_cf(P, it, ctl->item, block, RING ? ring.trips : nullptr);
    ctl->cur_item = it;
    // ctl->ent was last read by the consumers with ordinary loads (generic proxy); the bulk copy below writes it through the
    // async proxy.  The mbarrier hand-off (item_empty) orders the two only within the generic proxy: a proxy fence is required.
    fence_proxy_async();
    mbar_arrive_expect_tx(item_full, (uint32_t)sizeof(BlockTabEntry));
    tma_bulk_g2s(&ctl->ent, P.blocktab + block, (uint32_t)sizeof(BlockTabEntry), item_full);
    // the energy stage of this item (one item ahead of the consumers) reads 18 OV2 tiles of 2 KB: pull them into L2 now
    const int A = ctl->item.A, B = ctl->item.B, C = ctl->item.C;
    const int toff[3] = {ov2_tile_off(P, B, C), ov2_tile_off(P, A, B), ov2_tile_off(P, A, C)};
    const int occ[3] = {ctl->item.i, ctl->item.j, ctl->item.k};
#pragma unroll
    for (int x = 0; x < 3; x++)
#pragma unroll
        for (int y = 0; y < 3; y++)
            if (x != y) {
                const double* pb = P.OV2 + ov2_pair_base(P, occ[x], occ[y]);
#pragma unroll
                for (int t = 0; t < 3; t++) tma_prefetch_l2(pb + toff[t], 2048);
            }
}

// Producer: one thread.  Streams the Q chunks of every GEMM of every item through the ring.  Item n+1 is fetched and
// decoded while the chunks of item n are still streaming (right after its first GEMM), so the ring never runs dry at an
// item boundary.
// How an item updates the W slots.  0: accumulate path (RMW, twin-GEMM reuse).  Six-slot blocks: 1: load-accumulate-store over
// all 18 GEMMs (i > j > k); 2 / 3: i = j / j = k -- the 12 GEMMs of BlockTabEntry::sorder[mode - 2], W completed by symmetry in
// the energy stage (sym_emask).  dbg_flags 64 / 128 force the accumulate path for all items / for the symmetric classes.
__device__ __forceinline__ int item_mode(const Problem& P, const Ctl* ctl)
{
    if (!ctl->ent.fast_ok || (P.dbg_flags & 64)) return 0;
    if (ctl->item.i == ctl->item.j) return (P.dbg_flags & 128) ? 0 : 2;
    if (ctl->item.j == ctl->item.k) return (P.dbg_flags & 128) ? 0 : 3;
    return 1;
}
// t-th GEMM of an item in mode >= 1, and which of its column halves are kept / which are first contributions
__device__ __forceinline__ int seq_len(const Ctl* ctl, int mode) { return mode == 1 ? ctl->ent.ngemm : SYM_GEMMS; }
__device__ __forceinline__ int seq_gemm(const Ctl* ctl, int mode, int t) { return mode == 1 ? ctl->ent.forder[t] : ctl->ent.sorder[mode - 2][t]; }
__device__ __forceinline__ int seq_first(const Ctl* ctl, int mode, int t) { return mode == 1 ? ctl->ent.ffirst[t] : ctl->ent.sfirst[mode - 2][t]; }
__device__ __forceinline__ int seq_emask(const Ctl* ctl, int mode, int g) { return mode == 1 ? 3 : sym_emask(mode - 2, ctl->ent.gemm[g].p); }

template <bool FASTOK, bool RING, class Tail>
__device__ __forceinline__ void producer_loop(const Problem& P, const RingMap& ring, i64 item_begin, i64 item_end, unsigned long long* counter,
                                              double* Qsm, Tail* tail)
{
    const int nchunks = (P.G + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
    const i64 gstride = (i64)P.vp * KGROUP;   // doubles between consecutive kappa groups of one (q,r) in Qt
    int stage = 0;
    uint32_t sphase = 0;
    uint32_t nfetch = 0;
    mbar_wait((uint64_t*)&tail->item_empty[0], 1);
    producer_decode<RING>(P, ring, item_begin, item_end, counter, &tail->ctl[0], (uint64_t*)&tail->item_full[0], nfetch);
    for (uint32_t n = 0;; n++) {
        const int slot = n & 1;
        const Ctl* ctl = &tail->ctl[slot];
        mbar_wait((uint64_t*)&tail->item_full[slot], (n >> 1) & 1);   // own TMA copy of the table entry has landed
        if (ctl->cur_item < 0) break;
        const int mode = FASTOK ? item_mode(P, ctl) : 0;
        const int ngemm = mode ? seq_len(ctl, mode) : ctl->ent.ngemm;
        for (int t = 0; t < ngemm; t++) {
            const int g = mode ? seq_gemm(ctl, mode, t) : t;
            if (!mode && gemm_is_dup(ctl->item, g)) continue;   // no k-loop for twin GEMMs (i == j or j == k)
            const GemmDesc& gd = ctl->ent.gemm[g];
            const uint32_t row_bytes = (uint32_t)gd.TZ * KGROUP * sizeof(double);
            const int q = occ_pick(ctl->item, gd.q), r = occ_pick(ctl->item, gd.r);
            const double* src0 = P.Qt + qt_row(P, q, r, 0, gd.z0);
            const double* src1 = P.Qt + qt_row(P, r, q, 0, gd.z0);
            for (int c = 0; c < nchunks; c++) {
                const int ng = min(CHUNK_GROUPS, P.G - c * CHUNK_GROUPS);
                mbar_wait((uint64_t*)&tail->empty[stage], sphase ^ 1);
                // WAR across proxies: the consumers' LDS reads of this stage (generic proxy) must be ordered before the bulk
                // copies (async proxy) that refill it.  Acquiring `empty` is not enough -- with one-group stages the refill was
                // measurably overtaking the reads (wrong energies for 0.5 % of the 18-GEMM items) until this fence was added.
                fence_proxy_async();
                uint64_t* fb = (uint64_t*)&tail->full[stage];
                mbar_arrive_expect_tx(fb, (uint32_t)ng * 2u * row_bytes);
                double* st = Qsm + stage * QSTAGE_DOUBLES;
                for (int gl = 0; gl < ng; gl++) {
                    tma_bulk_g2s(st, src0, row_bytes, fb);
                    tma_bulk_g2s(st + QBLK, src1, row_bytes, fb);
                    st += 2 * QBLK;
                    src0 += gstride;
                    src1 += gstride;
                }
                if (++stage == QSTAGES) { stage = 0; sphase ^= 1; }
            }
            if (t == 0) {   // consumers are inside item n now, so ctl[slot^1] (item n-1) is, or soon will be, released
                const uint32_t m = n + 1;
                mbar_wait((uint64_t*)&tail->item_empty[m & 1], ((m >> 1) & 1) ^ 1);
                producer_decode<RING>(P, ring, item_begin, item_end, counter, &tail->ctl[m & 1], (uint64_t*)&tail->item_full[m & 1], nfetch);
            }
        }
        // Energy stage of this item: the ring area is reused for the 12 OV2 tiles whose row index is a, so that the a-loop of the
        // energy stage reads shared memory only.  Stage by stage, oldest first: as soon as the consumers have released a stage
        // (its last Q chunk is consumed) its four tiles are fetched, while the last k-loop still runs out of the other stages.
        {
            int st = stage;
            uint32_t ph = sphase;
            uint64_t* ob = (uint64_t*)&tail->ov_full;
            mbar_arrive_expect_tx(ob, (uint32_t)(OV_STAGE_TILES * 256 * sizeof(double)));
            for (int t = 0; t < QSTAGES; t++) {
                mbar_wait((uint64_t*)&tail->empty[st], ph ^ 1);
                fence_proxy_async();   // same hazard as above: the stage was read with LDS, the tiles arrive through the async proxy
#pragma unroll 1
                for (int u = 0; u < OV_TILES_PER_STAGE; u++)
                    tma_bulk_g2s(Qsm + st * QSTAGE_DOUBLES + u * 256, P.OV2 + ov2_stage_src(P, ctl->item, st * OV_TILES_PER_STAGE + u),
                                 256 * sizeof(double), ob);
                if (++st == QSTAGES) { st = 0; ph ^= 1; }
            }
            mbar_wait((uint64_t*)&tail->ov_empty, n & 1);   // energy stage done: the ring may be refilled
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// consumer: row pointers + first APREF A-fragment groups of a GEMM (issued before the previous GEMM's RMW epilogue so
// that their latency hides behind it)
// ---------------------------------------------------------------------------------------------------
struct RowSet {
    const double* base;        // P row of (x0, y0) for this p, + 2*kk   (warp-uniform part + lane kk)
    int off[MTW_MAX];          // per row-tile offset in doubles
    int nvalid;                // number of valid row tiles of this warp (0..mtw)
    int rt0;                   // first row tile of this warp
};

// Row tiles (8 rows) of a GEMM are dealt to the 16 consumer warps as evenly as possible: warp w owns `nv` consecutive tiles
// starting at rt0.  (ceil(rt_total/16) tiles for every warp would make the surplus warps multiply padding: 7 % of the DMMAs at C4.)
__device__ __forceinline__ void warp_rows(int rt_total, int w, int& rt0, int& nv)
{
    const int base = rt_total >> 4, extra = rt_total & (NCWARPS - 1);
    nv = base + (w < extra ? 1 : 0);
    rt0 = w * base + min(w, extra);
}

template <bool RING>
__device__ __forceinline__ void rows_setup2(const Problem& P, const RingMap& ring, const GemmDesc& gd, int p_orb, int warp, int lane, RowSet& rs)
{
    const int r = lane >> 2, kk = lane & 3;
    rs.base = P.Pt + pt_row(P, RING ? ring.pslot[p_orb] : p_orb, gd.y0, gd.x0) + 4 * kk;
    warp_rows(gd.rt_total, warp, rs.rt0, rs.nvalid);
#pragma unroll
    for (int mt = 0; mt < MTW_MAX; mt++) {
        const int m = (mt < rs.nvalid) ? (rs.rt0 + mt) * 8 + r : r;
        const int yl = (m * gd.xinv) >> 16;
        const int xl = m - yl * gd.TX;
        rs.off[mt] = (yl * P.vp + xl) * P.Kp;
    }
}

__device__ __forceinline__ void a_prologue2(const Problem& P, const RowSet& rs, double4x (&a)[ABUF][MTW_MAX])
{
    if (P.G > 0) {
#pragma unroll
        for (int mt = 0; mt < MTW_MAX; mt++)
            if (mt < rs.nvalid) a[0][mt] = ldg_stream_f64x4(rs.base + rs.off[mt]);
    }
}

// a warp without rows in this GEMM still takes part in the Q ring protocol
template <class Tail>
__device__ __forceinline__ void kloop_idle(const Problem& P, Tail* tail, int& stage, uint32_t& sphase, int lane)
{
    const int nchunks = (P.G + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
    for (int c = 0; c < nchunks; c++) {
        mbar_wait((uint64_t*)&tail->full[stage], sphase);
        __syncwarp();
        if (lane == 0) mbar_arrive((uint64_t*)&tail->empty[stage]);
        if (++stage == QSTAGES) { stage = 0; sphase ^= 1; }
    }
}

// accumulators -> TMEM, column order (e, mt, ct) so that the epilogue fetches the 4 column tiles of one (e, mt) with one x8 load
template <int MTW, int NT>
__device__ __forceinline__ void park_acc(const double (&acc)[MTW][NT][2], uint32_t taddr)
{
    uint32_t v[32];
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                const int idx = (e * 2 + mt) * 4 + ct;
                if (mt < MTW && ct < NT) {
                    v[2 * idx] = (uint32_t)__double2loint(acc[mt < MTW ? mt : 0][ct < NT ? ct : 0][e]);
                    v[2 * idx + 1] = (uint32_t)__double2hiint(acc[mt < MTW ? mt : 0][ct < NT ? ct : 0][e]);
                } else {
                    v[2 * idx] = 0u;
                    v[2 * idx + 1] = 0u;
                }
            }
    tmem_st32(taddr, v);
    tmem_wait_st();
}

// ---------------------------------------------------------------------------------------------------
// consumer: k-loop of one GEMM.  acc[mt][ct][e]: row tile mt, column tile ct, D element e.
// Column n of tile ct is (zl = 4ct + (n>>1), s = n&1), so a lane's two D elements are (zl = 4ct + kk, s = e).
// ---------------------------------------------------------------------------------------------------
template <int MTW, int NT, bool PROF, bool ZERO = true, class Tail>
__device__ __forceinline__ void gemm_kloop(const Problem& P, const GemmDesc& gd, const RowSet& rs, double4x (&a)[ABUF][MTW_MAX],
                                           double (&acc)[MTW][NT][2], const double* Qsm, Tail* tail, int& stage,
                                           uint32_t& sphase, int lane, long long* prof)
{
    const int kk = lane & 3, n = lane >> 2;
    const int boff = ((n & 1) * QBLK) + (n >> 1) * KGROUP + 4 * kk;   // s block + row zl(ct=0) + this lane's 4 kappa
    const int nchunks = (P.G + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
    if (ZERO) {
#pragma unroll
        for (int mt = 0; mt < MTW; mt++)
#pragma unroll
            for (int ct = 0; ct < NT; ct++) acc[mt][ct][0] = acc[mt][ct][1] = 0.0;
    }

    for (int c = 0; c < nchunks; c++) {
        const int g0 = c * CHUNK_GROUPS;
        const int ng = min(CHUNK_GROUPS, P.G - g0);
        long long tf = 0;
        if (PROF) tf = clock64();
        mbar_wait((uint64_t*)&tail->full[stage], sphase);
        if (PROF) prof[8] += clock64() - tf;
        const double* st = Qsm + stage * QSTAGE_DOUBLES + boff;
        int nstage = stage + 1;
        uint32_t nphase = sphase;
        if (nstage == QSTAGES) { nstage = 0; nphase ^= 1; }
#pragma unroll
        for (int gl = 0; gl < CHUNK_GROUPS; gl++) {
            if (gl < ng) {
                const int gg = g0 + gl;
                if (gg + APREF < P.G) {
#pragma unroll
                    for (int mt = 0; mt < MTW; mt++)
                        a[(gl + APREF) % ABUF][mt] = ldg_stream_f64x4(rs.base + rs.off[mt] + (gg + APREF) * KGROUP);
                }
                // kappa = 16*gg + 4*kk + h, h = 0..3: two halves of B fragments (h = 0,1 then h = 2,3)
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    double2 b[NT];
#pragma unroll
                    for (int ct = 0; ct < NT; ct++)
                        b[ct] = *reinterpret_cast<const double2*>(st + gl * 2 * QBLK + ct * 4 * KGROUP + 2 * hh);
#pragma unroll
                    for (int mt = 0; mt < MTW; mt++)
#pragma unroll
                        for (int ct = 0; ct < NT; ct++)
                            dmma884(acc[mt][ct][0], acc[mt][ct][1], hh ? a[gl % ABUF][mt].z : a[gl % ABUF][mt].x, b[ct].x);
#pragma unroll
                    for (int mt = 0; mt < MTW; mt++)
#pragma unroll
                        for (int ct = 0; ct < NT; ct++)
                            dmma884(acc[mt][ct][0], acc[mt][ct][1], hh ? a[gl % ABUF][mt].w : a[gl % ABUF][mt].y, b[ct].y);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive((uint64_t*)&tail->empty[stage]);
        stage = nstage;
        sphase = nphase;
    }
}

// The accumulators of one GEMM <-> its two W slots.  MODE 0: add into the slots (store where this is the slot's first
// contribution); MODE 1: load the slot contents into the accumulators (zero for a first contribution, bit s of `first_bits`);
// MODE 2: store the accumulators.
template <int MTW, int NT, int MODE>
__device__ __forceinline__ void gemm_wslots(const GemmDesc& gd, const RowSet& rs, double (&acc)[MTW][NT][2], double* Wsm, int lane,
                                            int first_bits, int emask = 3)
{
    const int kk = lane & 3, r = lane >> 2;
    int xl[MTW], yl[MTW];
#pragma unroll
    for (int mt = 0; mt < MTW; mt++) {
        const int m = (rs.rt0 + mt) * 8 + r;
        yl[mt] = (m * gd.xinv) >> 16;   // m / TX
        xl[mt] = m - yl[mt] * gd.TX;
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
        if (MODE != 0 && !((emask >> e) & 1)) {   // symmetric class: this column half is not kept
            if (MODE == 1) {
#pragma unroll
                for (int mt = 0; mt < MTW; mt++)
#pragma unroll
                    for (int ct = 0; ct < NT; ct++) acc[mt][ct][e] = 0.0;
            }
            continue;
        }
        // X and Z the same tile: D(s=0)[x=u,z=w] and D(s=1)[x=w,z=u] of different warps alias -> separate the column sets
        if (MODE == 0 && e == 1 && gd.diag_xz) consumer_bar();
        const int dbase = gd.dbase[e], sel = gd.dsel[e], Tb = gd.dTb[e], Tc = gd.dTc[e];
        const bool first = MODE == 0 ? gd.dfirst[e] != 0 : ((first_bits >> e) & 1) != 0;
#pragma unroll
        for (int mt = 0; mt < MTW; mt++) {
            {   // MTW is this warp's own tile count (dispatch on rs.nvalid): every mt < MTW is a valid row tile
                DestIter it;
                dest_iter_init_fast(dbase, sel, Tb, Tc, xl[mt], yl[mt], kk, it);
                if (MODE == 1) {
                    if (first) {
#pragma unroll
                        for (int ct = 0; ct < NT; ct++) acc[mt][ct][e] = 0.0;
                    } else {
#pragma unroll
                        for (int ct = 0; ct < NT; ct++) acc[mt][ct][e] = Wsm[dest_iter_off(it, ct)];
                    }
                } else if (MODE == 2 || first) {   // first contribution to this slot in the item: the slots are never zeroed
#pragma unroll
                    for (int ct = 0; ct < NT; ct++) Wsm[dest_iter_off(it, ct)] = acc[mt][ct][e];
                } else {
#pragma unroll
                    for (int ct = 0; ct < NT; ct++) Wsm[dest_iter_off(it, ct)] += acc[mt][ct][e];
                }
            }
        }
    }
}

// one GEMM of an item: k-loop, then (overlapped with the RMW epilogue) the next GEMM's row setup and first A loads
template <int MTW, int NT, bool PROF, bool RING>
__device__ __forceinline__ void gemm_body(const Problem& P, const RingMap& ring, const Ctl* ctl, int g, RowSet& rs, double4x (&a)[ABUF][MTW_MAX],
                                          double* Wsm, const double* Qsm, SmemTail* tail, int& stage, uint32_t& sphase,
                                          uint32_t& gcount, int warp, int lane, long long* prof)
{
    const GemmDesc& gd = ctl->ent.gemm[g];
    double acc[MTW > 0 ? MTW : 1][NT][2];
    long long t0 = 0, t1 = 0;
    if (PROF) t0 = clock64();
    if constexpr (MTW > 0) gemm_kloop<MTW, NT, PROF>(P, gd, rs, a, acc, Qsm, tail, stage, sphase, lane, prof);
    else kloop_idle(P, tail, stage, sphase, lane);   // no rows in this GEMM: only keep the ring protocol in step
    if (PROF) { t1 = clock64(); prof[2] += t1 - t0; }
    const RowSet rs_cur = rs;
    const bool dup_next = (g + 1 < ctl->ent.ngemm) && gemm_is_dup(ctl->item, g + 1);   // twin GEMM: same D, other destinations
    const int gnext = g + (dup_next ? 2 : 1);
    if (gnext < ctl->ent.ngemm) {
        const GemmDesc& gn = ctl->ent.gemm[gnext];
        rows_setup2<RING>(P, ring, gn, occ_pick(ctl->item, gn.p), warp, lane, rs);
        a_prologue2(P, rs, a);
    }
    for (int rep = 0; rep <= (dup_next ? 1 : 0); rep++) {
        long long tw = 0;
        if (PROF) tw = clock64();
        if (gcount > 0)   // every warp has finished the previous slot update (event gcount-1)
            mbar_wait((uint64_t*)&tail->rmw_done[(gcount - 1) & 1], ((gcount - 1) >> 1) & 1);
        if (PROF) { const long long t2 = clock64(); prof[6] += t2 - tw; tw = t2; }
        if constexpr (MTW > 0) {
            if (!(P.dbg_flags & 1)) gemm_wslots<MTW, NT, 0>(ctl->ent.gemm[g + rep], rs_cur, acc, Wsm, lane, 0);
            else if (acc[0][0][0] == 1.2345e300) Wsm[0] = acc[0][0][1];   // keep the accumulators alive
        } else if (ctl->ent.gemm[g + rep].diag_xz && !(P.dbg_flags & 1)) {
            consumer_bar();   // the CTA-wide barrier between the two column halves inside gemm_wslots
        }
        __syncwarp();
        if (PROF) prof[7] += clock64() - tw;
        if (lane == 0) mbar_arrive((uint64_t*)&tail->rmw_done[gcount & 1]);
        gcount++;
    }
    if (PROF) prof[3] += clock64() - t1;
}

// load-accumulate-store form of one GEMM (see the comment at the top): g = forder[t], gnext = forder[t+1] or -1
template <int MTW, int NT, bool PROF, bool RING>
__device__ __forceinline__ void gemm_body_fast(const Problem& P, const RingMap& ring, const Ctl* ctl, int g, int gnext, int first_bits, int emask, RowSet& rs,
                                               double4x (&a)[ABUF][MTW_MAX], double* Wsm, const double* Qsm, SmemTail* tail,
                                               int& stage, uint32_t& sphase, uint32_t& gcount, int warp, int lane, long long* prof)
{
    const GemmDesc& gd = ctl->ent.gemm[g];
    double acc[MTW > 0 ? MTW : 1][NT][2];
    long long t0 = 0, t1 = 0;
    if (PROF) t0 = clock64();
    if (gcount > 1)   // every warp's stores of the slot update two steps back (event gcount-2) are in the slots
        mbar_wait((uint64_t*)&tail->rmw_done[gcount & 1], ((gcount - 2) >> 1) & 1);
    if (PROF) { t1 = clock64(); prof[6] += t1 - t0; }
    if constexpr (MTW > 0) {
        if (!(P.dbg_flags & 1)) gemm_wslots<MTW, NT, 1>(gd, rs, acc, Wsm, lane, first_bits, emask);
        else gemm_wslots<MTW, NT, 1>(gd, rs, acc, Wsm, lane, 3, emask);
    }
    if (PROF) { t0 = clock64(); prof[7] += t0 - t1; }
    if constexpr (MTW > 0) gemm_kloop<MTW, NT, PROF, false>(P, gd, rs, a, acc, Qsm, tail, stage, sphase, lane, prof);
    else kloop_idle(P, tail, stage, sphase, lane);
    if (PROF) { t1 = clock64(); prof[2] += t1 - t0; }
    const RowSet rs_cur = rs;
    if (gnext >= 0) {
        const GemmDesc& gn = ctl->ent.gemm[gnext];
        rows_setup2<RING>(P, ring, gn, occ_pick(ctl->item, gn.p), warp, lane, rs);
        a_prologue2(P, rs, a);
    }
    if constexpr (MTW > 0) {
        if (!(P.dbg_flags & 1)) gemm_wslots<MTW, NT, 2>(gd, rs_cur, acc, Wsm, lane, 0, emask);
        else if (acc[0][0][0] == 1.2345e300) Wsm[0] = acc[0][0][1];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive((uint64_t*)&tail->rmw_done[gcount & 1]);
    gcount++;
    if (PROF) { const long long t2 = clock64(); prof[7] += t2 - t1; prof[3] += t2 - t1; }
}

#define FPT_DISPATCH_NT(MTWc, NTv, CALL)                                   \
    switch (NTv) {                                                         \
    case 1: { constexpr int MTW = MTWc, NT = 1; CALL; } break;             \
    case 2: { constexpr int MTW = MTWc, NT = 2; CALL; } break;             \
    case 3: { constexpr int MTW = MTWc, NT = 3; CALL; } break;             \
    default: { constexpr int MTW = MTWc, NT = 4; CALL; } break;            \
    }
#define FPT_DISPATCH(MTWv, NTv, CALL)                                      \
    do {                                                                   \
        switch (MTWv) {                                                    \
        case 0: { constexpr int MTW = 0, NT = 1; CALL; } break;            \
        case 1: FPT_DISPATCH_NT(1, NTv, CALL) break;                       \
        default: FPT_DISPATCH_NT(2, NTv, CALL) break;                      \
        }                                                                  \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// The kernel.  PROF: record a phase breakdown (cycles, warp 0's view) into prof_out[blockIdx*6 + 0..5] =
// {wait-for-item, zero, k-loops, RMW (+token wait, next prologue), energy, total, token wait, -} for warp 0 (entries 0-7)
// and for the first warp of the last group (entries 8-15).
// ---------------------------------------------------------------------------------------------------
template <bool PROF, bool RING = false>
__global__ void __launch_bounds__(NTHREADS, 1)
triples_kernel(Problem P, i64 item_begin, i64 item_end, unsigned long long* counter, double* partials, long long* prof_out, RingMap ring)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Wsm = reinterpret_cast<double*>(smem_raw);
    double* Qsm = Wsm + WSLOT_DOUBLES;
    SmemTail* tail = reinterpret_cast<SmemTail*>(Qsm + QSTAGES * QSTAGE_DOUBLES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < QSTAGES; s++) { mbar_init((uint64_t*)&tail->full[s], 1); mbar_init((uint64_t*)&tail->empty[s], NCWARPS); }
        for (int s = 0; s < 2; s++) { mbar_init((uint64_t*)&tail->item_full[s], 1); mbar_init((uint64_t*)&tail->item_empty[s], NCWARPS); }
        mbar_init((uint64_t*)&tail->rmw_done[0], NCWARPS);
        mbar_init((uint64_t*)&tail->rmw_done[1], NCWARPS);
        mbar_init((uint64_t*)&tail->ov_full, 1);
        mbar_init((uint64_t*)&tail->ov_empty, NCWARPS);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp >= NCWARPS) {   // producer warpgroup
        setmaxnreg_dec<PRODUCER_REGS>();
        if (warp == NCWARPS && lane == 0) producer_loop<true, RING>(P, ring, item_begin, item_end, counter, Qsm, tail);
        return;
    }
    setmaxnreg_inc<CONSUMER_REGS>();

    // ------------------------------- consumers -------------------------------
    long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_start = 0;
    if (PROF) t_start = clock64();
    double esum = 0.0;
    int stage = 0;
    uint32_t sphase = 0, gcount = 0;
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
        const int mode = item_mode(P, ctl);
        const int gfirst = mode ? seq_gemm(ctl, mode, 0) : 0;
        RowSet rs;
        rows_setup2<RING>(P, ring, ctl->ent.gemm[gfirst], occ_pick(ctl->item, ctl->ent.gemm[gfirst].p), warp, lane, rs);
        a_prologue2(P, rs, a);
        // the W slots are not zeroed: the first GEMM that reaches a slot stores into it (GemmDesc::dfirst / ffirst / sfirst);
        // the barrier at the end of the previous item's energy stage already ordered those stores after its reads
        if (PROF) { t0 = clock64(); prof[1] += t0 - t1; }

        if (mode) {
            const int nseq = seq_len(ctl, mode);
            for (int t = 0; t < nseq; t++) {
                const int g = seq_gemm(ctl, mode, t);
                const int gnext = t + 1 < nseq ? seq_gemm(ctl, mode, t + 1) : -1;
                const int fbits = seq_first(ctl, mode, t);
                const int emask = seq_emask(ctl, mode, g);
                const int nt = ctl->ent.gemm[g].TZ >> 2;
                // rs (set up during the previous GEMM) holds this warp's share of row tiles: 0, 1 or 2.  The broadcast tells the
                // compiler that the value is warp-uniform: without it the k-loops are compiled as potentially divergent code
                // (BSSY / WARPSYNC / BRA.DIV in the hot loop, no uniform-datapath instructions: -1.5 % on full-tile shapes)
                const int nv = __shfl_sync(0xffffffffu, rs.nvalid, 0);
                FPT_DISPATCH(nv, nt, (gemm_body_fast<MTW, NT, PROF, RING>(P, ring, ctl, g, gnext, fbits, emask, rs, a, Wsm, Qsm, tail, stage, sphase, gcount, warp, lane, prof)));
            }
        } else {
            for (int g = 0; g < ngemm; g++) {
                if (gemm_is_dup(ctl->item, g)) continue;   // handled by its twin (second RMW in gemm_body)
                const int nt = ctl->ent.gemm[g].TZ >> 2;
                const int nv = __shfl_sync(0xffffffffu, rs.nvalid, 0);   // warp-uniform, see above
                FPT_DISPATCH(nv, nt, (gemm_body<MTW, NT, PROF, RING>(P, ring, ctl, g, rs, a, Wsm, Qsm, tail, stage, sphase, gcount, warp, lane, prof)));
            }
        }
        if (PROF) t1 = clock64();
        consumer_bar();
        if (PROF) { t0 = clock64(); prof[10] += t0 - t1; }
        {
            const BlockDesc& bd = ctl->ent.bd;
            const int TC = bd.ts[2];
            const int half = tid >> 8, tt = tid & 255;   // two threads per (b,c) column, 8 values of a each
            mbar_wait((uint64_t*)&tail->ov_full, n & 1);      // the 12 a-row OV2 tiles are in the ring area
            if (PROF) prof[9] += clock64() - t0;
            if (!(P.dbg_flags & 2)) {
                if (bd.slot_elems == 4096)
                    esum += block_column_energy_t<true>(P, bd, ctl->item.i, ctl->item.j, ctl->item.k, Wsm, Qsm, tt >> 4, tt & 15,
                                                        half * 8, half * 8 + 8, mode >= 2 ? mode - 1 : 0);
                else if (tt < bd.ts[1] * TC)
                    esum += block_column_energy_t<false>(P, bd, ctl->item.i, ctl->item.j, ctl->item.k, Wsm, Qsm, tt / TC, tt % TC,
                                                         half * 8, half * 8 + 8, mode >= 2 ? mode - 1 : 0);
            }
        }
        if (PROF) t1 = clock64();
        consumer_bar();       // W slots, the staged tiles and ctl[slot] may be reused
        if (PROF) prof[11] += clock64() - t1;
        if (lane == 0) {
            mbar_arrive((uint64_t*)&tail->ov_empty);
            mbar_arrive((uint64_t*)&tail->item_empty[slot]);
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
    if (PROF && lane == 0 && (warp == 0 || warp == NCWARPS - 4)) {   // group 0's and group 3's view
        prof[5] = clock64() - t_start;
        for (int t = 0; t < 12; t++) prof_out[blockIdx.x * NPROF + (warp ? 12 : 0) + t] = prof[t];
    }
}

}  // namespace fpt
