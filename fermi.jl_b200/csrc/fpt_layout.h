// Index algebra of the B200 (T) path, shared by the CUDA kernels (fpt_kernels.cuh) and by the host-side
// emulator used in the CPU tests (tests/emul/emul_main.cpp).  Everything here is __host__ __device__ and
// free of CUDA runtime calls.
//
// Math being laid out (reference: src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:108-136 and the
// GEMM form ijk2.jl:26-33,110-153).  For occupied p and an ordered pair (q,r):
//
//     X(p;q,r)[x,y,z] = sum_kappa  P_p[(x,y),kappa] * Q_qr[kappa,z]
//     P_p[(x,y),kappa] = OVVV[p,y,x,d]   (kappa = d < v)      |  -T2[p,l,y,x]   (kappa = v+l)
//     Q_qr[kappa,z]    = T2[r,q,z,d]     (kappa = d < v)      |  OOOV[l,q,r,z]  (kappa = v+l)
//
//     W_ijk[a,b,c] = X(j;i,k)[a,b,c] + X(k;i,j)[a,c,b] + X(i;j,k)[b,a,c]
//                  + X(k;j,i)[b,c,a] + X(i;k,j)[c,a,b] + X(j;k,i)[c,b,a]
//
// i.e. X(p;q,r)[x,y,z] lands on the W element whose i/j/k-paired virtual is  x<->q, y<->p, z<->r.
//
// Device-resident operand layouts (built once per call by the prep kernels):
//     Pt [p][y][x][kappa]          kappa contiguous, Kp = roundup16(v+o) doubles per row, x,y < vp (zero padded)
//     Qt [q*o+r][g][z][16]         kappa = 16g + (0..15), z < vp (zero padded)
//     OV2[q*o+r][Y][Z][16][16] = OVOV[q,y,r,z] in 16x16 tiles (Y = y>>4, ...), zero padded;   T1d[p][x] = T1[p,x]
//
// Work decomposition: virtual range [0,vp) is cut into tiles (edge 16, last tile 4/8/12); a *block* is a
// tile triple A>=B>=C; an *item* is (i>=j, block, k<=j) ordered pair-major / block / k-fastest so that
// concurrently running CTAs share the P_i and P_j rows in L2.  One CTA owns one item at a time: it keeps
// W_ijk on all distinct permutations of the tile triple (<= 6 slots of TA*TB*TC doubles) in shared memory,
// accumulates 3*nslot P-stationary GEMMs into them and then evaluates ijk.jl:120-136 for the a>=b>=c
// points of the block without W or V ever leaving the SM.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FPT_HD __host__ __device__ __forceinline__
#else
#define FPT_HD inline
#endif

namespace fpt {

constexpr int TMAX = 16;       // largest tile edge
constexpr int KGROUP = 16;     // kappa per group: one 32-byte (256-bit) load per lane feeds four DMMA.8x8x4
constexpr int CHUNK_GROUPS = 2;  // kappa groups per shared-memory Q stage (32 kappa)
constexpr int MAX_SLOTS = 6;
constexpr int MAX_GEMMS = 18;
constexpr int SYM_GEMMS = 12;   // GEMMs of a six-slot block that run for a triplet with i = j or j = k

typedef long long i64;

FPT_HD int roundup(int x, int m) { return (x + m - 1) / m * m; }

// ---- tiles ---------------------------------------------------------------------------------------
// vp = roundup4(v) is cut into nt = ceil(vp/16) tiles: 16, 16, ..., then a remainder tile of 8 or 12; a remainder of 4
// is merged with the last full tile and re-split as 12 + 8 (a 4-wide tile makes N = 8 GEMMs that are bound by the A
// stream).  All tile sizes are multiples of 4.
FPT_HD int padded_v(int v) { return roundup(v, 4); }
FPT_HD int num_tiles(int v) { return (padded_v(v) + TMAX - 1) / TMAX; }
FPT_HD bool split_12_8(int vp) { return (vp & 15) == 4 && vp > 16; }
FPT_HD int tile_start(int t, int vp)
{
    const int nt = (vp + TMAX - 1) / TMAX;
    if (split_12_8(vp) && t == nt - 1) return 16 * (nt - 2) + 12;
    return t * TMAX;
}
FPT_HD int tile_size(int t, int vp)
{
    const int nt = (vp + TMAX - 1) / TMAX;
    if (split_12_8(vp)) return t < nt - 2 ? 16 : (t == nt - 2 ? 12 : 8);
    const int s = vp - t * TMAX;
    return s > TMAX ? TMAX : s;
}
FPT_HD int tile_of(int y, int vp)
{
    const int nt = (vp + TMAX - 1) / TMAX;
    if (split_12_8(vp) && y >= 16 * (nt - 2) + 12) return nt - 1;
    return y >> 4;
}

// ---- tetrahedral / triangular decodes ----------------------------------------------------------------
// n -> (A,B,C), A>=B>=C>=0, n = A(A+1)(A+2)/6 + B(B+1)/2 + C
FPT_HD void tetra_decode(i64 n, int& A, int& B, int& C)
{
    int a = 0;
    while ((i64)(a + 1) * (a + 2) * (a + 3) / 6 <= n) a++;
    n -= (i64)a * (a + 1) * (a + 2) / 6;
    int b = 0;
    while ((i64)(b + 1) * (b + 2) / 2 <= n) b++;
    n -= (i64)b * (b + 1) / 2;
    A = a; B = b; C = (int)n;
}
FPT_HD void tri_decode(int n, int& i, int& j)
{
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= n) a++;
    i = a; j = n - a * (a + 1) / 2;
}
FPT_HD i64 num_blocks(int nt) { return (i64)nt * (nt + 1) * (nt + 2) / 6; }
// k values for pair (i,j): k in [0, j], minus the zero-weight i=j=k triplet (ijk.jl:133)
FPT_HD int num_k(int i, int j) { return (i == j) ? j : j + 1; }

// ---- shared-memory W slots ---------------------------------------------------------------------------
// Element (la,lb,lc) of a slot with dims (Ta,Tb,Tc), lc fastest, with a bank swizzle on lc chosen so that
//  (1) a 4x4 patch in any two coordinates (lo-2-bits aligned) and (2) a run of 16 in any single coordinate
// hit 16 distinct 8-byte banks (RMW epilogue of the GEMMs, and the energy stage's permuted reads).
// For Tc == 16 the swizzle is GF(2)-linear:  lcs = lc ^ SA(la) ^ SB(lb),  SA(x) = swap the two bit pairs of x,
// SB(x) = (x&3)*5 ^ (x>>2); linearity is what makes the per-column-tile offsets of the RMW epilogue one XOR away
// from each other (DestIter below).  Edge tiles (Tc = 4/8/12) swizzle only the low two bits of lc.
FPT_HD int swz_a(int x) { return ((x & 3) << 2) | (x >> 2); }
FPT_HD int swz_b(int x) { return ((x & 3) * 5) ^ (x >> 2); }
FPT_HD int slot_index(int la, int lb, int lc, int Tb, int Tc)
{
    // Tc == 16: full 4-bit XOR swizzle; edge tiles (4/8/12): only the low two bits are swizzled (stays inside the tile)
    const int mask = (Tc == 16) ? 15 : 3;
    return (la * Tb + lb) * Tc + (lc ^ ((swz_a(la) ^ swz_b(lb)) & mask));
}

// perm m of three things, m = 0..5: (0,1,2),(0,2,1),(1,0,2),(1,2,0),(2,0,1),(2,1,0); perm3(m,c) = c-th entry
FPT_HD int perm3(int m, int c)
{
    const int f = m >> 1;
    if (c == 0) return f;
    const int lo = (f == 0) ? 1 : 0, hi = (f == 2) ? 1 : 2;
    return ((c == 1) == ((m & 1) == 0)) ? lo : hi;
}

struct BlockDesc {
    int tile[3];       // tile indices (A,B,C), A>=B>=C
    int t0[3];         // tile starts
    int ts[3];         // tile sizes
    int nslot;         // distinct permuted tile triples
    int slot_of_perm[6];   // perm m of (A,B,C) -> compact slot id
    int perm_of_slot[6];   // representative perm for each slot
    int slot_elems;        // TA*TB*TC
};

FPT_HD void make_block(int A, int B, int C, int vp, BlockDesc& bd)
{
    bd.tile[0] = A; bd.tile[1] = B; bd.tile[2] = C;
    for (int c = 0; c < 3; c++) { bd.t0[c] = tile_start(bd.tile[c], vp); bd.ts[c] = tile_size(bd.tile[c], vp); }
    bd.slot_elems = bd.ts[0] * bd.ts[1] * bd.ts[2];
    bd.nslot = 0;
    for (int m = 0; m < 6; m++) {
        int found = -1;
        for (int m2 = 0; m2 < m; m2++) {
            bool same = true;
            for (int c = 0; c < 3; c++) same = same && (bd.tile[perm3(m, c)] == bd.tile[perm3(m2, c)]);
            if (same) { found = bd.slot_of_perm[m2]; break; }
        }
        if (found < 0) { bd.perm_of_slot[bd.nslot] = m; found = bd.nslot++; }
        bd.slot_of_perm[m] = found;
    }
}

// One P-stationary GEMM:  D[(x,y), (s,z)] = sum_kappa P_p[(x,y),kappa] * Q_{s}[kappa,z],
//   s=0: Q_{q r}, s=1: Q_{r q};  x in tile X (rows x fastest), y in tile Y, z in tile Z.
struct GemmDesc {
    int p, q, r;            // occupied orbital numbers
    int x0, y0, z0;         // tile starts of X, Y, Z
    int TX, TY, TZ;         // tile sizes
    // destination of D[...,s,...]: slot base (doubles), which of (x,y,z) supplies (la,lb,lc), and dest Tb,Tc
    int dbase[2];
    int dsel[2];            // packed: ra | rb<<2 | rc<<4   (0=x,1=y,2=z)
    int dTb[2], dTc[2];
    int diag_xz;            // X and Z are the same tile: D(s=0) and D(s=1) of different threads alias in the W slot
    int dfirst[2];          // D[...,s,...] is the first contribution its W slot receives in this item: store, do not add
    int xinv;               // ceil(65536 / TX): row m = yl*TX + xl  ->  yl = (m * xinv) >> 16 for m < 256
    int rt_total;           // TX*TY/8 row tiles of 8
};

// occ = (i,j,k).  Builds the 3*nslot GEMMs of an item.  Returns their number.
FPT_HD int make_gemms(const BlockDesc& bd, int i, int j, int k, GemmDesc* gd)
{
    const int occ[3] = {i, j, k};
    int n = 0;
    bool touched[MAX_SLOTS] = {false, false, false, false, false, false};
    for (int sl = 0; sl < bd.nslot; sl++) {
        const int m = bd.perm_of_slot[sl];
        const int cx = perm3(m, 0), cy = perm3(m, 1), cz = perm3(m, 2);  // which of A/B/C plays X, Y, Z
        for (int pi = 0; pi < 3; pi++) {
            GemmDesc& g = gd[n++];
            const int qi = (pi == 0) ? 1 : 0, ri = (pi == 2) ? 1 : 2;  // the two other positions, ascending
            g.p = occ[pi]; g.q = occ[qi]; g.r = occ[ri];
            g.x0 = bd.t0[cx]; g.y0 = bd.t0[cy]; g.z0 = bd.t0[cz];
            g.TX = bd.ts[cx]; g.TY = bd.ts[cy]; g.TZ = bd.ts[cz];
            g.diag_xz = (bd.tile[cx] == bd.tile[cz]);
            g.xinv = (65536 + g.TX - 1) / g.TX;
            g.rt_total = (g.TX * g.TY) >> 3;
            for (int s = 0; s < 2; s++) {
                // pairing: x <-> (s ? r : q), y <-> p, z <-> (s ? q : r).  sel[pos] = which of x/y/z pairs with occ pos
                int sel[3];
                sel[pi] = 1;
                sel[s ? ri : qi] = 0;
                sel[s ? qi : ri] = 2;
                // destination tile triple in terms of A/B/C positions: (c_of[sel[0]], c_of[sel[1]], c_of[sel[2]])
                const int cxyz[3] = {cx, cy, cz};
                const int da = cxyz[sel[0]], db = cxyz[sel[1]], dc = cxyz[sel[2]];
                // find the perm producing the same tile triple
                int dslot = -1;
                for (int m2 = 0; m2 < 6 && dslot < 0; m2++)
                    if (bd.tile[perm3(m2, 0)] == bd.tile[da] && bd.tile[perm3(m2, 1)] == bd.tile[db] &&
                        bd.tile[perm3(m2, 2)] == bd.tile[dc])
                        dslot = bd.slot_of_perm[m2];
                g.dbase[s] = dslot * bd.slot_elems;
                g.dsel[s] = sel[0] | (sel[1] << 2) | (sel[2] << 4);
                g.dTb[s] = bd.ts[db];
                g.dTc[s] = bd.ts[dc];
                // GEMMs run in index order (a twin GEMM's second add directly follows its twin), so "first" is static
                g.dfirst[s] = touched[dslot] ? 0 : 1;
                touched[dslot] = true;
            }
        }
    }
    return n;
}

FPT_HD int pick3(int sel, int x, int y, int z) { return sel == 0 ? x : (sel == 1 ? y : z); }

// smem offset (doubles) of the destination of D element (xl,yl,zl) for column set s
FPT_HD int gemm_dest(const GemmDesc& g, int s, int xl, int yl, int zl)
{
    const int sel = g.dsel[s];
    const int la = pick3(sel & 3, xl, yl, zl), lb = pick3((sel >> 2) & 3, xl, yl, zl), lc = pick3((sel >> 4) & 3, xl, yl, zl);
    return g.dbase[s] + slot_index(la, lb, lc, g.dTb[s], g.dTc[s]);
}

// Fast form of gemm_dest for the RMW epilogue: for a fixed thread (xl, yl, kk) the element for column tile ct
// (zl = 4*ct + kk) is
//     off(ct) = lin0 + ct*zs + (w0 ^ (ct*d1))
// where (zs, d1) = (0, 4) if z supplies lc, (4*Tc, 1) if z supplies lb, (4*Tb*Tc, 1) if z supplies la
// (swz_a(4ct) = swz_b(4ct) = ct, 4ct + kk = 4ct ^ kk, and ct < 4 only touches the low two bits).
struct DestIter { int lin0, zs, w0, d1; };

// `sel` (which of x,y,z supplies la,lb,lc) is uniform per (GEMM, s): switch on it once, no per-element selects remain.
FPT_HD void dest_iter_init_fast(int dbase, int sel, int Tb, int Tc, int xl, int yl, int kk, DestIter& it)
{
    int la, lb, lc;
    switch (sel) {
    case (0 | (1 << 2) | (2 << 4)): la = xl; lb = yl; lc = kk; it.zs = 0; it.d1 = 4; break;            // (x,y,z)
    case (0 | (2 << 2) | (1 << 4)): la = xl; lb = kk; lc = yl; it.zs = 4 * Tc; it.d1 = 1; break;       // (x,z,y)
    case (1 | (0 << 2) | (2 << 4)): la = yl; lb = xl; lc = kk; it.zs = 0; it.d1 = 4; break;            // (y,x,z)
    case (1 | (2 << 2) | (0 << 4)): la = yl; lb = kk; lc = xl; it.zs = 4 * Tc; it.d1 = 1; break;       // (y,z,x)
    case (2 | (0 << 2) | (1 << 4)): la = kk; lb = xl; lc = yl; it.zs = 4 * Tb * Tc; it.d1 = 1; break;  // (z,x,y)
    default:                         la = kk; lb = yl; lc = xl; it.zs = 4 * Tb * Tc; it.d1 = 1; break;  // (z,y,x)
    }
    const int mask = (Tc == 16) ? 15 : 3;
    it.lin0 = dbase + (la * Tb + lb) * Tc;
    it.w0 = lc ^ ((swz_a(la) ^ swz_b(lb)) & mask);
}
FPT_HD int dest_iter_off(const DestIter& it, int ct) { return it.lin0 + ct * it.zs + (it.w0 ^ (ct * it.d1)); }

// Per-block descriptors are independent of (i,j,k) up to which occupied plays p, q, r: the host builds one entry per
// block with make_gemms(bd, 0, 1, 2), so that GemmDesc.p/q/r hold *positions* in (i,j,k); the kernel's producer copies
// the entry of an item's block into shared memory with one TMA bulk copy.
struct BlockTabEntry {
    BlockDesc bd;
    int ngemm;
    GemmDesc gemm[MAX_GEMMS];
    // "Load-accumulate-store" order for blocks of three distinct tiles (nslot == 6): a permutation of the 18 GEMMs in which
    // two consecutive GEMMs never write the same W slot.  A GEMM can then start from the current slot contents (accumulators
    // initialised by LDS, or zero where ffirst says it is the slot's first contribution in this order) and finish with
    // plain stores: the only ordering it needs is against the GEMM *two* steps back, a whole k-loop earlier.
    unsigned char forder[MAX_GEMMS + 2];
    unsigned char ffirst[MAX_GEMMS + 2];   // bit s: D[..., s, ...] of forder[t] is the first contribution to its slot
    int fast_ok;
    int pad_;
    // The same for triplets with two equal occupied indices (class 0: i = j, class 1: j = k), where only 12 GEMMs run and the
    // W tensor is completed by symmetry in the energy stage (see sym_emask): order and first-contribution flags of those 12.
    unsigned char sorder[2][SYM_GEMMS];
    unsigned char sfirst[2][SYM_GEMMS];
};
static_assert(sizeof(BlockTabEntry) % 16 == 0, "BlockTabEntry is moved by 16-byte-granular bulk copies");

FPT_HD int gemm_slot(const BlockDesc& bd, const GemmDesc& g, int s) { return g.dbase[s] / bd.slot_elems; }

// Triplets with i = j (class 0) or j = k (class 1) in a six-slot block.  For i = j the operands of the p = j GEMM equal those
// of the p = i GEMM and the two column halves of the p = k GEMM are equal (Q_ij = Q_ji); the contributions that are *not*
// computed are exactly the computed ones with the first two W indices exchanged:  W[x,y,z] = T[x,y,z] + T[y,x,z], where T
// collects  p = i (both halves)  and  p = k (half s = 0 only).  For j = k:  W[x,y,z] = T[x,y,z] + T[x,z,y]  with T from
// p = j (both halves) and p = i (half s = 0).  The energy stage adds the mirror image when it reads the slots
// (block_column_energy_t, `sym`), so 12 of the 18 GEMMs run, none is added twice, and every one of them can take the
// load-accumulate-store form.  sym_emask: which column halves of GEMM position `pi` (0 = i, 1 = j, 2 = k) are kept
// (bit s), 0 = the GEMM does not run.
FPT_HD int sym_emask(int cls, int pi)
{
    if (cls == 0) return pi == 0 ? 3 : (pi == 2 ? 1 : 0);
    return pi == 1 ? 3 : (pi == 0 ? 1 : 0);
}

// depth-first search for an order of `n` nodes (GEMM index, kept halves) in which neighbours write different slots
// (18 or 12 nodes, dense compatibility graph: trivial); host only
inline bool slot_order_dfs(const BlockTabEntry& e, const unsigned char* node_g, const unsigned char* node_mask, int n,
                           unsigned char* order, bool* used, int depth)
{
    if (depth == n) return true;
    for (int c = 0; c < n; c++) {
        if (used[c]) continue;
        if (depth > 0) {
            const int pr = order[depth - 1];
            bool clash = false;
            for (int s = 0; s < 2; s++)
                for (int t = 0; t < 2; t++)
                    if (((node_mask[pr] >> s) & 1) && ((node_mask[c] >> t) & 1))
                        clash = clash || gemm_slot(e.bd, e.gemm[node_g[pr]], s) == gemm_slot(e.bd, e.gemm[node_g[c]], t);
            if (clash) continue;
        }
        used[c] = true; order[depth] = (unsigned char)c;
        if (slot_order_dfs(e, node_g, node_mask, n, order, used, depth + 1)) return true;
        used[c] = false;
    }
    return false;
}

inline void make_fast_order(BlockTabEntry& e)
{
    e.fast_ok = 0; e.pad_ = 0;
    for (int t = 0; t < MAX_GEMMS + 2; t++) { e.forder[t] = 0; e.ffirst[t] = 0; }
    for (int c = 0; c < 2; c++)
        for (int t = 0; t < SYM_GEMMS; t++) { e.sorder[c][t] = 0; e.sfirst[c][t] = 0; }
    if (e.bd.nslot != MAX_SLOTS || e.ngemm != MAX_GEMMS) return;
    // all 18 GEMMs, both halves (triplets i > j > k)
    {
        unsigned char ng[MAX_GEMMS], nm[MAX_GEMMS], ord[MAX_GEMMS];
        bool used[MAX_GEMMS] = {};
        for (int g = 0; g < MAX_GEMMS; g++) { ng[g] = (unsigned char)g; nm[g] = 3; }
        if (!slot_order_dfs(e, ng, nm, MAX_GEMMS, ord, used, 0)) return;
        bool touched[MAX_SLOTS] = {};
        for (int t = 0; t < MAX_GEMMS; t++) {
            e.forder[t] = ng[ord[t]];
            for (int s = 0; s < 2; s++) {
                const int sl = gemm_slot(e.bd, e.gemm[e.forder[t]], s);
                if (!touched[sl]) { e.ffirst[t] |= (unsigned char)(1 << s); touched[sl] = true; }
            }
        }
    }
    // the 12 GEMMs of the two symmetric classes
    for (int c = 0; c < 2; c++) {
        unsigned char ng[SYM_GEMMS], nm[SYM_GEMMS], ord[SYM_GEMMS];
        bool used[SYM_GEMMS] = {};
        int n = 0;
        for (int g = 0; g < MAX_GEMMS; g++) {
            const int m = sym_emask(c, e.gemm[g].p);   // table entries hold positions 0/1/2 in p
            if (m) { ng[n] = (unsigned char)g; nm[n] = (unsigned char)m; n++; }
        }
        if (n != SYM_GEMMS || !slot_order_dfs(e, ng, nm, SYM_GEMMS, ord, used, 0)) return;
        bool touched[MAX_SLOTS] = {};
        for (int t = 0; t < SYM_GEMMS; t++) {
            e.sorder[c][t] = ng[ord[t]];
            for (int s = 0; s < 2; s++) {
                if (!((nm[ord[t]] >> s) & 1)) continue;
                const int sl = gemm_slot(e.bd, e.gemm[e.sorder[c][t]], s);
                if (!touched[sl]) { e.sfirst[c][t] |= (unsigned char)(1 << s); touched[sl] = true; }
            }
        }
        for (int sl = 0; sl < MAX_SLOTS; sl++)
            if (!touched[sl]) return;   // every slot must receive something (it does: three halves per slot)
    }
    e.fast_ok = 1;
}

// Relative cost of one item of a block, in units of one DMMA.8x8x4 on one SM sub-partition, for the static cost-weighted
// split of the work list across GPUs: per GEMM and kappa group the busiest sub-partition issues (row tiles of its four
// warps) x (TZ/4) x 4 DMMAs (row tiles are dealt evenly to the 16 warps, warp w sits on sub-partition w % 4), plus a fixed
// epilogue (slot update, next GEMM's setup) worth about 2 full-size groups; per item the energy stage and the item switch
// cost about 5 full-size groups (phase profile profiles/r01_phase_c4_mid.json, tuned on the measured 8-way shard times at C4).
FPT_HD double block_cost(const BlockTabEntry& e, int G)
{
    double c = 5.0 * 128.0;
    for (int g = 0; g < e.ngemm; g++) {
        const int rt = e.gemm[g].rt_total, nt = e.gemm[g].TZ >> 2;
        const int tiles = 4 * (rt >> 4) + (((rt & 15) + 3) >> 2);
        c += (double)(tiles * nt * 4) * G + 2.0 * 128.0;
    }
    return c;
}

// ---- problem description --------------------------------------------------------------------------
struct Problem {
    int o, v, vp, nt, Kp, G;
    i64 nb;        // blocks per triplet
    // Work list: items = (non-zero-weight triplet u in the window [tw_begin, tw_begin + tw_count)) x (block).
    // order 1 (default): block-major, u fastest -- every CTA of the grid works on the same tile triple, so the
    //     6*o P panels of that block (42 MB at C4) stay L2-resident while all triplets stream through them;
    // order 0: triplet-major, block fastest -- concurrent CTAs share P_i, P_j, P_k of one triplet.
    int order;
    i64 tw_begin, tw_count;
    i64 nitems;    // nb * tw_count
    const double* Pt;
    const double* Qt;
    const double* OV2;
    const double* T1d;
    const double* fo;
    const double* fv;
    const BlockTabEntry* blocktab;   // nb entries (device)
    int dbg_flags;            // diagnostics only (results become wrong): 1 = skip RMW epilogues, 2 = skip energy stage
};

FPT_HD i64 pt_row(const Problem& P, int p, int y, int x) { return (((i64)p * P.vp + y) * P.vp + x) * P.Kp; }
// Slab-ring mode of the density-fitted route (fpt_api.cu, triples_df_ring): Pt holds only a few occupied slabs at a time --
// pslot[p] = slab position of occupied p (nullptr: Pt holds all o slabs, position p) -- and the triplets of a launch are an explicit
// list, trips[3u .. 3u+2] = (i, j, k) of entry u (nullptr: entry u of the reference's list, decoded arithmetically).  Kept out of
// `Problem`: the fused kernel's default instantiation must not change with it (its k-loops are sensitive to every register).
struct RingMap { const int* pslot; const int* trips; };
FPT_HD i64 pt_row_slot(const Problem& P, const int* pslot, int p, int y, int x) { return pt_row(P, pslot ? pslot[p] : p, y, x); }
FPT_HD i64 qt_row(const Problem& P, int q, int r, int g, int z) { return ((((i64)q * P.o + r) * P.G + g) * P.vp + z) * KGROUP; }

struct ItemDesc { int i, j, k; int A, B, C; };

// OV2 is stored in 16x16 tiles so that the 18 tiles an item's energy stage reads are 2 KB contiguous blocks (one TMA L2
// prefetch each) and the in-tile offsets are shift/adds.
FPT_HD i64 ov2_pair_base(const Problem& P, int q, int r) { return ((i64)q * P.o + r) * P.nt * P.nt * 256; }
FPT_HD int ov2_tile_off(const Problem& P, int Y, int Z) { return (Y * P.nt + Z) * 256; }
FPT_HD i64 ov2_idx(const Problem& P, int q, int r, int y, int z)
{
    const int Y = tile_of(y, P.vp), Z = tile_of(z, P.vp);
    return ov2_pair_base(P, q, r) + ov2_tile_off(P, Y, Z) + ((y - tile_start(Y, P.vp)) << 4) + (z - tile_start(Z, P.vp));
}
FPT_HD i64 ov2_elems(const Problem& P) { return (i64)P.o * P.o * P.nt * P.nt * 256; }

FPT_HD int occ_pick(const ItemDesc& it, int pos) { return pos == 0 ? it.i : (pos == 1 ? it.j : it.k); }

// The GEMMs of an item come in triples (same operand tiles, p = i, j, k).  For i == j the p = j GEMM has exactly the
// operands of the p = i one (P_i = P_j, Q_jk = Q_ik, Q_kj = Q_ki), for j == k the p = k GEMM repeats the p = j one: such
// a GEMM is not recomputed, the accumulators of its twin are added a second time at the twin's destinations.
FPT_HD bool gemm_is_dup(const ItemDesc& it, int g)
{
    const int pi = g % 3;
    return (pi == 1 && it.i == it.j) || (pi == 2 && it.j == it.k);
}

// Non-zero-weight triplets i >= j >= k (not i = j = k, ijk.jl:133) in the reference's loop order (i, then j, then k
// fastest; ijk.jl:49,63,83) are numbered u = 0 .. ntrip-1; T(i,j) = i(i+1)(i+2)/6 - i + j(j+1)/2 of them precede pair (i,j).
FPT_HD i64 num_triplets(int o) { return (i64)o * (o + 1) * (o + 2) / 6 - o; }
FPT_HD void triplet_decode(int o, i64 u, int& i_, int& j_, int& k_)
{
    int i = 0;
    while (i + 1 < o && (i64)(i + 1) * (i + 2) * (i + 3) / 6 - (i + 1) <= u) i++;
    const i64 ti = (i64)i * (i + 1) * (i + 2) / 6 - i;
    int j = 0;   // pair (i,j) holds num_k(i,j) = j+1 triplets (j < i) or i (j = i, the last pair of this i)
    while (j + 1 <= i && (i64)(j + 1) * (j + 2) / 2 <= u - ti) j++;
    i_ = i; j_ = j; k_ = (int)(u - ti - (i64)j * (j + 1) / 2);
}
// position t in the reference's full list of i >= j >= k triplets (zero-weight ones included) -> number of
// non-zero-weight triplets before it (the diagonal triplet (m,m,m) sits at t = m(m+1)(m+2)/6 + m(m+1)/2 + m)
FPT_HD i64 triplets_before(int o, i64 t)
{
    i64 ndiag = 0;
    for (int m = 0; m < o; m++)
        if ((i64)m * (m + 1) * (m + 2) / 6 + (i64)m * (m + 1) / 2 + m < t) ndiag++;
    return t - ndiag;
}

FPT_HD void item_decode_cf(const Problem& P, i64 item, ItemDesc& it, i64& block, const int* trips = nullptr)
{
    i64 u;
    if (P.order == 1) { block = item / P.tw_count; u = item - block * P.tw_count; }
    else              { u = item / P.nb; block = item - u * P.nb; }
    if (trips) { it.i = trips[3 * u]; it.j = trips[3 * u + 1]; it.k = trips[3 * u + 2]; }
    else triplet_decode(P.o, P.tw_begin + u, it.i, it.j, it.k);
    tetra_decode(block, it.A, it.B, it.C);
}

// Static split of the item range [b, e) into `world` contiguous parts of equal estimated cost; part `rank` is [*sb, *se).
// In block-major order an item's cost depends on its block (diagonal and edge blocks are cheaper), so boundaries are placed on
// the prefix sum of block_cost (one entry per block); in triplet-major order every stretch of nb items costs the same.
// `frac` (optional, world + 1 values rising from 0 to 1): boundary r sits at that fraction of the range's estimated cost instead of
// r / world -- how the adaptive balance of a multi-GPU handle shifts the shards (fpt_api_compute.inl).
inline void shard_items(const Problem& P, const double* cost, i64 b, i64 e, int rank, int world, i64* sb, i64* se, const double* frac = nullptr)
{
    if (P.order != 1 || P.tw_count <= 0) {
        *sb = b + (e - b) * rank / world;
        *se = b + (e - b) * (rank + 1) / world;
        return;
    }
    // cost of items [0, x): whole blocks plus a partial one
    auto cost_upto = [&](i64 x) {
        const i64 blk = x / P.tw_count, rem = x - blk * P.tw_count;
        double c = 0.0;
        for (i64 t = 0; t < blk; t++) c += cost[t] * (double)P.tw_count;
        if (blk < P.nb) c += cost[blk] * (double)rem;
        return c;
    };
    const double c0 = cost_upto(b), c1 = cost_upto(e);
    auto boundary = [&](int r) -> i64 {
        if (r <= 0) return b;
        if (r >= world) return e;
        const double target = c0 + (c1 - c0) * (frac ? frac[r] : (double)r / world);
        double c = 0.0;
        for (i64 blk = 0; blk < P.nb; blk++) {
            const double cb = cost[blk] * (double)P.tw_count;
            if (c + cb >= target) {
                i64 x = blk * P.tw_count + (i64)((target - c) / cost[blk] + 0.5);
                return x < b ? b : (x > e ? e : x);
            }
            c += cb;
        }
        return e;
    };
    *sb = boundary(rank);
    *se = boundary(rank + 1);
}

// Adaptive balance: `frac` (W + 1 boundary fractions, 0 = frac[0] < ... < frac[W] = 1) gave the shards the measured times ms[0..W).
// With the time density of old segment r taken as ms[r] / (frac[r+1] - frac[r]), boundary k belongs where the cumulative time reaches
// k T / W; the move towards it is damped.  Returns false (and leaves `out` unspecified) if the result would not be strictly rising.
inline bool rebalance_fractions(int W, const double* frac, const double* ms, double damping, double* out)
{
    double T = 0.0;
    for (int r = 0; r < W; r++) T += ms[r];
    out[0] = 0.0;
    out[W] = 1.0;
    int r = 0;
    double acc = 0.0;   // time of the segments before r
    for (int k = 1; k < W; k++) {
        const double target = T * k / W;
        while (r < W - 1 && acc + ms[r] < target) acc += ms[r++];
        const double w = frac[r + 1] - frac[r];
        const double u = ms[r] > 0.0 ? frac[r] + w * (target - acc) / ms[r] : frac[r];
        out[k] = frac[k] + damping * (u - frac[k]);
    }
    for (int k = 0; k < W; k++)
        if (!(out[k + 1] > out[k])) return false;
    return true;
}

// ---- energy of one (a,b,c) point, ijk.jl:127-133 ---------------------------------------------------------
// w[m], vv[m] in perm order m: (abc),(acb),(bac),(bca),(cab),(cba)
FPT_HD double point_energy(const double* w, const double* vv, double Dd, int a, int b, int c, double wijk)
{
    const double X = w[0] * vv[0] + w[1] * vv[1] + w[2] * vv[2] + w[3] * vv[3] + w[4] * vv[4] + w[5] * vv[5];
    const double Y = vv[0] + vv[3] + vv[4];
    const double Z = vv[1] + vv[2] + vv[5];
    const double Ef = (Y - 2.0 * Z) * (w[0] + w[3] + w[4]) + (Z - 2.0 * Y) * (w[1] + w[2] + w[5]) + 3.0 * X;
    const double den = Dd * (double)(1 + (a == b) + (b == c));
    return Ef * wijk / den;
}

// Energy contribution of point `pt` (linear index, c fastest) of the block held in the W slots `Wsm`:
// V build ijk.jl:116 and the a>=b>=c body ijk.jl:127-133.  Returns 0 for padded or non-canonical points.
FPT_HD double block_point_energy(const Problem& P, const BlockDesc& bd, int i, int j, int k, const double* Wsm, int pt, int sym = 0)
{
    const int TB = bd.ts[1], TC = bd.ts[2];
    const int v = P.v;
    const int cl = pt % TC;
    const int t2 = pt / TC;
    const int bl = t2 % TB, al = t2 / TB;
    const int a = bd.t0[0] + al, b = bd.t0[1] + bl, c = bd.t0[2] + cl;
    if (a >= v || b >= v || c >= v || a < b || b < c) return 0.0;
    const double* t1i = P.T1d + (i64)i * v;
    const double* t1j = P.T1d + (i64)j * v;
    const double* t1k = P.T1d + (i64)k * v;
    // OV2[(q,r)][y][z] = (qy|rz) = OV2[(r,q)][z][y]: always index so that the virtual that comes later in (a,b,c)
    // is the contiguous one -- lanes run over c, so every load is either coalesced or a broadcast.
    double w[6], vv[6];
    for (int m = 0; m < 6; m++) {
        const int c0 = perm3(m, 0), c1 = perm3(m, 1), c2 = perm3(m, 2);
        const int lx = pick3(c0, al, bl, cl), ly = pick3(c1, al, bl, cl), lz = pick3(c2, al, bl, cl);
        w[m] = Wsm[bd.slot_of_perm[m] * bd.slot_elems + slot_index(lx, ly, lz, bd.ts[c1], bd.ts[c2])];
    }
    if (sym == 1) {          // i = j: W[x,y,z] = T[x,y,z] + T[y,x,z]   (perm order: abc, acb, bac, bca, cab, cba)
        const double s02 = w[0] + w[2], s14 = w[1] + w[4], s35 = w[3] + w[5];
        w[0] = w[2] = s02; w[1] = w[4] = s14; w[3] = w[5] = s35;
    } else if (sym == 2) {   // j = k: W[x,y,z] = T[x,y,z] + T[x,z,y]
        const double s01 = w[0] + w[1], s23 = w[2] + w[3], s45 = w[4] + w[5];
        w[0] = w[1] = s01; w[2] = w[3] = s23; w[4] = w[5] = s45;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < 6; m++) {
        const int c0 = perm3(m, 0), c1 = perm3(m, 1), c2 = perm3(m, 2);
        const int x = pick3(c0, a, b, c), y = pick3(c1, a, b, c), z = pick3(c2, a, b, c);
        const double g_jk = (c2 > c1) ? P.OV2[ov2_idx(P, j, k, y, z)] : P.OV2[ov2_idx(P, k, j, z, y)];   // (jy|kz)
        const double g_ik = (c2 > c0) ? P.OV2[ov2_idx(P, i, k, x, z)] : P.OV2[ov2_idx(P, k, i, z, x)];   // (ix|kz)
        const double g_ij = (c1 > c0) ? P.OV2[ov2_idx(P, i, j, x, y)] : P.OV2[ov2_idx(P, j, i, y, x)];   // (ix|jy)
        vv[m] = w[m] + t1i[x] * g_jk + g_ik * t1j[y] + g_ij * t1k[z];
    }
    const double Dd = P.fo[i] + P.fo[j] + P.fo[k] - P.fv[a] - P.fv[b] - P.fv[c];
    return point_energy(w, vv, Dd, a, b, c, (double)(2 - (i == j) - (j == k)));
}

// Staging area of the energy stage's OV2 tiles: tile t = 2*pair + col sits in ring stage t/4 at tile slot t%4, so that the
// producer can refill a ring stage with its four tiles as soon as the consumers have released that stage, while the last
// k-loop still runs out of the other stages.  OV_STAGE_STRIDE = doubles per ring stage (checked in fpt_triples.cuh).
constexpr int OV_STAGE_STRIDE = 1096;
constexpr int OV_TILES_PER_STAGE = 4;
FPT_HD int ov_stage_off(int t) { return (t / OV_TILES_PER_STAGE) * OV_STAGE_STRIDE + (t % OV_TILES_PER_STAGE) * 256; }
#define OVP(k) ((((k) >> 1) * fpt::OV_STAGE_STRIDE) + (((k) & 1) * 512))   /* offset of pair k's (B, C) tile couple */

// `ovs` holds the 12 OV2 tiles whose row index is a -- [(pair: jk,kj,ik,ki,ij,ji) x (column tile: B, C)][16][16] -- staged in
// shared memory by the producer (TMA) so that the a-loop has no global loads; see ov2_stage_src.
template <bool ALL16>
FPT_HD double block_column_energy_t(const Problem& P, const BlockDesc& bd, int i, int j, int k, const double* Wsm,
                                    const double* ovs, int bl, int cl, int al_begin, int al_end, int sym = 0)
{
    const int TA = ALL16 ? 16 : bd.ts[0], TB = ALL16 ? 16 : bd.ts[1], TC = ALL16 ? 16 : bd.ts[2];
    const int v = P.v;
    const int a0 = bd.t0[0];
    const int b = bd.t0[1] + bl, c = bd.t0[2] + cl;
    if (b >= v || c >= v || b < c) return 0.0;
    // a runs over [max(al_begin, b - a0), min(al_end, TA, v - a0)): a >= b (ijk.jl:123) and a < v (padding)
    if (al_begin < b - a0) al_begin = b - a0;
    if (al_end > TA) al_end = TA;
    if (al_end > v - a0) al_end = v - a0;
    if (al_begin >= al_end) return 0.0;
    const double* t1i = P.T1d + (i64)i * v;
    const double* t1j = P.T1d + (i64)j * v;
    const double* t1k = P.T1d + (i64)k * v;
    const double* ovjk = P.OV2 + ov2_pair_base(P, j, k);   // [y][z] = (jy|kz)
    const double* ovkj = P.OV2 + ov2_pair_base(P, k, j);   // [y][z] = (ky|jz) = (jz|ky)
    const double* ovik = P.OV2 + ov2_pair_base(P, i, k);
    const double* ovki = P.OV2 + ov2_pair_base(P, k, i);
    const double* ovij = P.OV2 + ov2_pair_base(P, i, j);
    const double* ovji = P.OV2 + ov2_pair_base(P, j, i);
    // tile-local offsets: (y,z) in tile (Y,Z) sits at tile_off + yl*16 + zl
    const int bc = ov2_tile_off(P, bd.tile[1], bd.tile[2]) + (bl << 4) + cl;
    const double* sab = ovs + bl;          // tile (pair, B): [al][bl]
    const double* sac = ovs + 256 + cl;    // tile (pair, C): [al][cl]
    const double jk_bc = ovjk[bc], jk_cb = ovkj[bc];       // (jb|kc), (jc|kb)
    const double ik_bc = ovik[bc], ik_cb = ovki[bc];       // (ib|kc), (ic|kb)
    const double ij_bc = ovij[bc], ij_cb = ovji[bc];       // (ib|jc), (ic|jb)
    const double t1i_b = t1i[b], t1j_b = t1j[b], t1k_b = t1k[b];
    const double t1i_c = t1i[c], t1j_c = t1j[c], t1k_c = t1k[c];
    const double Dbc = P.fo[i] + P.fo[j] + P.fo[k] - P.fv[b] - P.fv[c];
    const double wijk = (double)(2 - (i == j) - (j == k));
    const int se = bd.slot_elems;
    const int s0 = bd.slot_of_perm[0] * se, s1 = bd.slot_of_perm[1] * se, s2 = bd.slot_of_perm[2] * se;
    const int s3 = bd.slot_of_perm[3] * se, s4 = bd.slot_of_perm[4] * se, s5 = bd.slot_of_perm[5] * se;
    double e = 0.0;
    for (int al = al_begin; al < al_end; al++) {
        const int a = a0 + al;
        const int ab = al << 4, ac = al << 4;
        double w0 = Wsm[s0 + slot_index(al, bl, cl, TB, TC)];   // W[a,b,c]
        double w1 = Wsm[s1 + slot_index(al, cl, bl, TC, TB)];   // W[a,c,b]
        double w2 = Wsm[s2 + slot_index(bl, al, cl, TA, TC)];   // W[b,a,c]
        double w3 = Wsm[s3 + slot_index(bl, cl, al, TC, TA)];   // W[b,c,a]
        double w4 = Wsm[s4 + slot_index(cl, al, bl, TA, TB)];   // W[c,a,b]
        double w5 = Wsm[s5 + slot_index(cl, bl, al, TB, TA)];   // W[c,b,a]
        if (sym == 1) {          // i = j: the slots hold T, W[x,y,z] = T[x,y,z] + T[y,x,z] (sym_emask)
            const double s02 = w0 + w2, s14 = w1 + w4, s35 = w3 + w5;
            w0 = w2 = s02; w1 = w4 = s14; w3 = w5 = s35;
        } else if (sym == 2) {   // j = k: W[x,y,z] = T[x,y,z] + T[x,z,y]
            const double s01 = w0 + w1, s23 = w2 + w3, s45 = w4 + w5;
            w0 = w1 = s01; w2 = w3 = s23; w4 = w5 = s45;
        }
        // V = W + D with D the disconnected term (ijk.jl:116); every product t1*g is consumed as soon as it is formed:
        //   X = sum_m W_m V_m = Xw + Xd,  Y = V0+V3+V4 = Ye + Yd,  Z = V1+V2+V5 = Zo + Zd
        const double Ye = w0 + w3 + w4, Zo = w1 + w2 + w5;
        double X = w0 * w0 + w1 * w1 + w2 * w2 + w3 * w3 + w4 * w4 + w5 * w5;
        double Yd, Zd, p;
        const double t1i_a = t1i[a], t1j_a = t1j[a], t1k_a = t1k[a];
        // D0 = t1i_a jk_bc + ik_ac t1j_b + ij_ab t1k_c
        p = t1i_a * jk_bc + sac[OVP(2) + ac] * t1j_b + sab[OVP(4) + ab] * t1k_c;  Yd = p;  X += w0 * p;
        // D1 = t1i_a jk_cb + ik_ab t1j_c + ij_ac t1k_b
        p = t1i_a * jk_cb + sab[OVP(2) + ab] * t1j_c + sac[OVP(4) + ac] * t1k_b;  Zd = p;  X += w1 * p;
        // D2 = t1i_b jk_ac + ik_bc t1j_a + ij_ba t1k_c
        p = t1i_b * sac[OVP(0) + ac] + ik_bc * t1j_a + sab[OVP(5) + ab] * t1k_c;  Zd += p; X += w2 * p;
        // D3 = t1i_b jk_ca + ik_ba t1j_c + ij_bc t1k_a
        p = t1i_b * sac[OVP(1) + ac] + sab[OVP(3) + ab] * t1j_c + ij_bc * t1k_a;  Yd += p; X += w3 * p;
        // D4 = t1i_c jk_ab + ik_cb t1j_a + ij_ca t1k_b
        p = t1i_c * sab[OVP(0) + ab] + ik_cb * t1j_a + sac[OVP(5) + ac] * t1k_b;  Yd += p; X += w4 * p;
        // D5 = t1i_c jk_ba + ik_ca t1j_b + ij_cb t1k_a
        p = t1i_c * sab[OVP(1) + ab] + sac[OVP(3) + ac] * t1j_b + ij_cb * t1k_a;  Zd += p; X += w5 * p;
        const double Y = Ye + Yd, Z = Zo + Zd;
        const double Ef = (Y - 2.0 * Z) * Ye + (Z - 2.0 * Y) * Zo + 3.0 * X;                       // ijk.jl:132
        const double den = (Dbc - P.fv[a]) * (double)(1 + (a == b) + (b == c));                   // ijk.jl:133
        e += Ef * wijk / den;
    }
    return e;
}

// Energy of the column (bl, cl) of the block for al in [al_begin, al_end): same arithmetic as block_point_energy, organised
// so that one thread owns (b,c), hoists everything that does not depend on a and walks a with 12 loads per point, each
// either contiguous in c across the lanes or a broadcast.  This is what the kernel runs; the emulator checks it.
FPT_HD double block_column_energy(const Problem& P, const BlockDesc& bd, int i, int j, int k, const double* Wsm,
                                  const double* ovs, int bl, int cl, int al_begin, int al_end, int sym = 0)
{
    if (bd.ts[0] == 16 && bd.ts[1] == 16 && bd.ts[2] == 16)
        return block_column_energy_t<true>(P, bd, i, j, k, Wsm, ovs, bl, cl, al_begin, al_end, sym);
    return block_column_energy_t<false>(P, bd, i, j, k, Wsm, ovs, bl, cl, al_begin, al_end, sym);
}

// Source (offset into OV2) of staged tile t = 2*pair + col, pair in (jk,kj,ik,ki,ij,ji), col 0 = tile B, 1 = tile C,
// rows = tile A, for item (i,j,k) and block (A,B,C).
FPT_HD i64 ov2_stage_src(const Problem& P, const ItemDesc& it, int t)
{
    const int pair = t >> 1, col = t & 1;
    const int q = (pair == 0 || pair == 5) ? it.j : ((pair == 1 || pair == 3) ? it.k : it.i);
    const int r = (pair == 0 || pair == 2) ? it.k : ((pair == 1 || pair == 4) ? it.j : it.i);
    return ov2_pair_base(P, q, r) + ov2_tile_off(P, it.A, col ? it.C : it.B);
}
constexpr int OV_STAGE_TILES = 12;

}  // namespace fpt
