// Host-side emulator of the fused B200 (T) kernel's data flow (test infrastructure, CPU only).
//
// It includes the very same index algebra the CUDA kernels use (fermi.jl_b200/csrc/fpt_layout.h: item decode,
// block/slot construction, GEMM descriptors, swizzled slot addressing, the per-point energy) and replaces only
// the tensor-core GEMM and the thread mapping by plain loops.  tests/test_emulator.py builds it with g++ and
// checks its E(T) against the oracle, so layout bugs are caught without a GPU.
//
// Exported C function:
//   fpt_emulate(o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, order, t_begin, t_end, item_begin, item_end, &Et, &nitems)
// (order / triplet window / item range as in fpt_set_item_order, fpt_set_triplet_window, fpt_compute)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../fermi.jl_b200/csrc/fpt_layout.h"

using namespace fpt;

extern "C" int fpt_emulate(int o, int v, const double* T1, const double* T2, const double* OVVV, const double* OOOV,
                           const double* OVOV, const double* fo, const double* fv, int order, long long t_begin, long long t_end,
                           long long item_begin, long long item_end, double* Et, long long* nitems_out)
{
    Problem P{};
    P.o = o; P.v = v; P.vp = padded_v(v); P.nt = num_tiles(v);
    P.Kp = roundup(v + o, KGROUP); P.G = P.Kp / KGROUP;
    P.nb = num_blocks(P.nt);
    // independent enumeration of the reference's triplet list (ijk.jl:49,63,83) to check the closed-form decodes against
    struct Trip { int i, j, k; };
    std::vector<Trip> trips;        // non-zero-weight triplets of the window, in order
    {
        i64 t = 0;
        const i64 nfull = (i64)o * (o + 1) * (o + 2) / 6;
        if (t_end < 0 || t_end > nfull) t_end = nfull;
        for (int i = 0; i < o; i++)
            for (int j = 0; j <= i; j++)
                for (int k = 0; k <= j; k++, t++)
                    if (t >= t_begin && t < t_end && !(i == j && j == k)) trips.push_back({i, j, k});
    }
    P.order = order;
    P.tw_begin = triplets_before(o, t_begin);
    P.tw_count = triplets_before(o, t_end) - P.tw_begin;
    if (P.tw_count != (i64)trips.size()) { fprintf(stderr, "triplets_before disagrees with the enumeration\n"); return 9; }
    P.nitems = P.nb * P.tw_count;
    if (nitems_out) *nitems_out = P.nitems;

    // ---- layout prep (same formulas as prep_* kernels) ----
    std::vector<double> Pt((size_t)o * P.vp * P.vp * P.Kp, 0.0), Qt((size_t)o * o * P.G * P.vp * KGROUP, 0.0),
        OV2((size_t)ov2_elems(P), 0.0), T1d((size_t)o * v);
    for (int p = 0; p < o; p++)
        for (int y = 0; y < v; y++)
            for (int x = 0; x < v; x++) {
                double* row = Pt.data() + pt_row(P, p, y, x);
                for (int d = 0; d < v; d++) row[d] = OVVV[p + (i64)o * (y + (i64)v * (x + (i64)v * d))];
                for (int l = 0; l < o; l++) row[v + l] = -T2[p + (i64)o * (l + (i64)o * (y + (i64)v * x))];
            }
    for (int q = 0; q < o; q++)
        for (int r = 0; r < o; r++)
            for (int z = 0; z < v; z++)
                for (int kappa = 0; kappa < v + o; kappa++) {
                    double val = kappa < v ? T2[r + (i64)o * (q + (i64)o * (z + (i64)v * kappa))]
                                           : OOOV[(kappa - v) + (i64)o * (q + (i64)o * (r + (i64)o * z))];
                    Qt[qt_row(P, q, r, kappa / KGROUP, z) + (kappa % KGROUP)] = val;
                }
    for (int q = 0; q < o; q++)
        for (int r = 0; r < o; r++)
            for (int y = 0; y < v; y++)
                for (int z = 0; z < v; z++)
                    OV2[ov2_idx(P, q, r, y, z)] = OVOV[q + (i64)o * (y + (i64)v * (r + (i64)o * z))];
    for (int p = 0; p < o; p++)
        for (int x = 0; x < v; x++) T1d[(i64)p * v + x] = T1[p + (i64)o * x];
    P.Pt = Pt.data(); P.Qt = Qt.data(); P.OV2 = OV2.data(); P.T1d = T1d.data(); P.fo = fo; P.fv = fv;

    if (item_end < 0 || item_end > P.nitems) item_end = P.nitems;
    std::vector<double> W((size_t)MAX_SLOTS * TMAX * TMAX * TMAX);
    double E = 0.0;
    for (i64 item = item_begin; item < item_end; item++) {
        ItemDesc it;
        i64 blk;
        item_decode_cf(P, item, it, blk);
        {
            const i64 u = order == 1 ? item % P.tw_count : item / P.nb;
            const i64 b = order == 1 ? item / P.tw_count : item % P.nb;
            const Trip& tr = trips[(size_t)u];
            if (it.i != tr.i || it.j != tr.j || it.k != tr.k || blk != b) {
                fprintf(stderr, "closed-form item decode disagrees at %lld\n", item);
                return 7;
            }
        }
        if (!(it.i >= it.j && it.j >= it.k) || (it.i == it.j && it.j == it.k) || !(it.A >= it.B && it.B >= it.C) || it.A >= P.nt) {
            fprintf(stderr, "bad item decode %lld -> %d %d %d / %d %d %d\n", item, it.i, it.j, it.k, it.A, it.B, it.C);
            return 2;
        }
        BlockDesc bd;
        make_block(it.A, it.B, it.C, P.vp, bd);
        GemmDesc gd[MAX_GEMMS];
        const int ng = make_gemms(bd, 0, 1, 2, gd);   // positions, as in the device block table
        BlockTabEntry ent{};
        ent.bd = bd; ent.ngemm = ng;
        for (int g = 0; g < ng; g++) ent.gemm[g] = gd[g];
        make_fast_order(ent);
        for (int g = 0; g < ng; g++) { gd[g].p = occ_pick(it, gd[g].p); gd[g].q = occ_pick(it, gd[g].q); gd[g].r = occ_pick(it, gd[g].r); }
        {   // the load-accumulate-store order of the block table (BlockTabEntry::forder / ffirst) must be a permutation of the
            // GEMMs in which neighbours write different slots, with exactly the first contribution of every slot flagged
            if ((bd.nslot == MAX_SLOTS) != (ent.fast_ok != 0)) { fprintf(stderr, "fast order missing for a six-slot block\n"); return 10; }
            if (ent.fast_ok) {
                bool seen[MAX_GEMMS] = {}, touched[MAX_SLOTS] = {};
                for (int t = 0; t < ng; t++) {
                    const int g = ent.forder[t];
                    if (g >= ng || seen[g]) { fprintf(stderr, "fast order is not a permutation\n"); return 11; }
                    seen[g] = true;
                    for (int s2 = 0; s2 < 2; s2++) {
                        const int sl = gemm_slot(bd, gd[g], s2);
                        if (t > 0)
                            for (int s3 = 0; s3 < 2; s3++)
                                if (sl == gemm_slot(bd, gd[ent.forder[t - 1]], s3)) { fprintf(stderr, "neighbours share a slot\n"); return 12; }
                        const bool first = !touched[sl];
                        if (first != (((ent.ffirst[t] >> s2) & 1) != 0)) { fprintf(stderr, "wrong first-contribution flag\n"); return 13; }
                        touched[sl] = true;
                    }
                    if (gemm_slot(bd, gd[g], 0) == gemm_slot(bd, gd[g], 1)) { fprintf(stderr, "six-slot block with aliasing halves\n"); return 14; }
                }
            }
        }
        // triplets with i = j or j = k in a six-slot block: only the 12 GEMMs of BlockTabEntry::sorder run, half of them keep
        // one column half only, and the energy stage completes W by symmetry (sym_emask).  Check order, flags and energy.
        const int cls = (it.i == it.j) ? 0 : ((it.j == it.k) ? 1 : -1);
        const bool symmetric = ent.fast_ok && cls >= 0;
        if (symmetric) {
            bool seen[MAX_GEMMS] = {}, touched[MAX_SLOTS] = {};
            for (int t = 0; t < SYM_GEMMS; t++) {
                const int g = ent.sorder[cls][t];
                const int em = sym_emask(cls, ent.gemm[g].p);
                if (g >= ng || seen[g] || em == 0) { fprintf(stderr, "bad symmetric order\n"); return 15; }
                seen[g] = true;
                for (int s2 = 0; s2 < 2; s2++) {
                    if (!((em >> s2) & 1)) continue;
                    const int sl = gemm_slot(bd, gd[g], s2);
                    if (t > 0) {
                        const int gp = ent.sorder[cls][t - 1], emp = sym_emask(cls, ent.gemm[gp].p);
                        for (int s3 = 0; s3 < 2; s3++)
                            if (((emp >> s3) & 1) && sl == gemm_slot(bd, gd[gp], s3)) { fprintf(stderr, "symmetric neighbours share a slot\n"); return 16; }
                    }
                    const bool first = !touched[sl];
                    if (first != (((ent.sfirst[cls][t] >> s2) & 1) != 0)) { fprintf(stderr, "wrong symmetric first flag\n"); return 17; }
                    touched[sl] = true;
                }
            }
        }
        std::fill(W.begin(), W.begin() + (size_t)bd.nslot * bd.slot_elems, 0.0);
        std::vector<int> hits((size_t)bd.nslot * bd.slot_elems, 0);
        std::vector<double> Dbuf;   // D of the last computed GEMM, [m][s][zl]
        for (int g = 0; g < ng; g++) {
            const GemmDesc& G = gd[g];
            const int emask = symmetric ? sym_emask(cls, ent.gemm[g].p) : 3;
            if (emask == 0) continue;   // symmetric class: this GEMM does not run
            const bool dup = !symmetric && gemm_is_dup(it, g);   // twin GEMM: the kernel reuses the accumulators instead of recomputing
            if (!dup) Dbuf.assign((size_t)G.TX * G.TY * 2 * G.TZ, 0.0);
            for (int m = 0; m < G.TX * G.TY; m++) {
                const int yl = m / G.TX, xl = m % G.TX;
                const double* prow = P.Pt + pt_row(P, G.p, G.y0 + yl, G.x0 + xl);
                for (int s = 0; s < 2; s++)
                    for (int zl = 0; zl < G.TZ; zl++) {
                        if (!((emask >> s) & 1)) continue;   // symmetric class: this column half is not kept
                        double& d = Dbuf[((size_t)m * 2 + s) * G.TZ + zl];
                        if (!dup) {
                            d = 0.0;
                            for (int kappa = 0; kappa < P.Kp; kappa++)
                                d += prow[kappa] * P.Qt[qt_row(P, s ? G.r : G.q, s ? G.q : G.r, kappa / KGROUP, G.z0 + zl) + (kappa % KGROUP)];
                        }
                        const int off = gemm_dest(G, s, xl, yl, zl);
                        {   // the kernel's fast RMW addressing must agree with the reference form
                            DestIter df;
                            dest_iter_init_fast(G.dbase[s], G.dsel[s], G.dTb[s], G.dTc[s], xl, yl, zl & 3, df);
                            if (dest_iter_off(df, zl >> 2) != off) { fprintf(stderr, "dest_iter_fast mismatch\n"); return 8; }
                        }
                        if (off < 0 || off >= bd.nslot * bd.slot_elems) { fprintf(stderr, "dest out of range\n"); return 3; }
                        W[off] += d;
                        hits[off]++;
                    }
            }
        }
        for (size_t t = 0; t < hits.size(); t++)
            if (hits[t] != (symmetric ? 3 : 6)) { fprintf(stderr, "slot element %zu received %d contributions\n", t, hits[t]); return 4; }
        const int sym = symmetric ? cls + 1 : 0;
        std::vector<double> ovs((size_t)(OV_STAGE_TILES / OV_TILES_PER_STAGE) * OV_STAGE_STRIDE);
        for (int t = 0; t < OV_STAGE_TILES; t++)
            for (int e = 0; e < 256; e++) ovs[(size_t)ov_stage_off(t) + e] = P.OV2[ov2_stage_src(P, it, t) + e];
        double e_pt = 0.0, e_col = 0.0;
        for (int pt = 0; pt < bd.slot_elems; pt++) e_pt += block_point_energy(P, bd, it.i, it.j, it.k, W.data(), pt, sym);
        for (int bl = 0; bl < bd.ts[1]; bl++)
            for (int cl = 0; cl < bd.ts[2]; cl++) e_col += block_column_energy(P, bd, it.i, it.j, it.k, W.data(), ovs.data(), bl, cl, 0, 8, sym) +
                         block_column_energy(P, bd, it.i, it.j, it.k, W.data(), ovs.data(), bl, cl, 8, 16, sym);
        if (std::fabs(e_pt - e_col) > 1e-13 * (1e-30 + std::fabs(e_pt)) + 1e-18) {
            fprintf(stderr, "column energy %.17g != point energy %.17g\n", e_col, e_pt);
            return 6;
        }
        E += e_col;
    }
    *Et = E;
    return 0;
}

// The library's cost-weighted split of the work list (fpt_shard_items) for the CPU tests: same code (shard_items, block_cost,
// make_gemms of fpt_layout.h), no GPU.  Returns the estimated cost share of the part in *cost_share (1/world when balanced).
extern "C" int fpt_emul_shard_frac(int o, int v, int order, int rank, int world, const double* frac, long long* item_begin,
                                   long long* item_end, double* cost_share);
extern "C" int fpt_emul_shard(int o, int v, int order, int rank, int world, long long* item_begin, long long* item_end,
                              double* cost_share)
{
    return fpt_emul_shard_frac(o, v, order, rank, world, nullptr, item_begin, item_end, cost_share);
}
// the same with the boundary fractions of the adaptive balance (frac: world + 1 values, nullptr = uniform)
extern "C" int fpt_emul_shard_frac(int o, int v, int order, int rank, int world, const double* frac, long long* item_begin,
                                   long long* item_end, double* cost_share)
{
    Problem P{};
    P.o = o; P.v = v; P.vp = padded_v(v); P.nt = num_tiles(v);
    P.Kp = roundup(v + o, KGROUP); P.G = P.Kp / KGROUP;
    P.nb = num_blocks(P.nt);
    P.order = order; P.tw_begin = 0; P.tw_count = num_triplets(o); P.nitems = P.nb * P.tw_count;
    std::vector<double> cost((size_t)P.nb);
    for (i64 b = 0; b < P.nb; b++) {
        BlockTabEntry ent{};
        int A, B, C;
        tetra_decode(b, A, B, C);
        make_block(A, B, C, P.vp, ent.bd);
        ent.ngemm = make_gemms(ent.bd, 0, 1, 2, ent.gemm);
        cost[(size_t)b] = block_cost(ent, P.G);
    }
    i64 sb, se;
    shard_items(P, cost.data(), 0, P.nitems, rank, world, &sb, &se, frac);
    *item_begin = sb; *item_end = se;
    if (cost_share) {
        double tot = 0.0, mine = 0.0;
        for (i64 it = 0; it < P.nitems; it++) {
            const i64 blk = order == 1 ? it / P.tw_count : it % P.nb;
            tot += cost[(size_t)blk];
            if (it >= sb && it < se) mine += cost[(size_t)blk];
        }
        *cost_share = tot > 0 ? mine / tot : 0.0;
    }
    return 0;
}


// The adaptive balance's boundary update (rebalance_fractions in fpt_layout.h), for the CPU tests.
extern "C" int fpt_emul_rebalance(int W, const double* frac, const double* ms, double damping, double* out)
{
    return rebalance_fractions(W, frac, ms, damping, out) ? 0 : 1;
}
