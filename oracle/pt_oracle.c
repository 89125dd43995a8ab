/*
 * CPU ORACLE (test infrastructure, NOT product code) -- plain-C restatement of Fermi.jl's RCCSD(T).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product path (fermi.jl_b200 + libfermi_pt_b200.so) never does.
 *
 * Reference followed (paths relative to /root/reference):
 *   src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:20-150   default algorithm (pt_alg=1)
 *   src/Methods/CoupledCluster/PerturbativeTriples/ijk2.jl:21-190  same math as explicit GEMMs (pt_alg=2)
 *
 * Two entry points:
 *   pt_oracle_naive  literal scalar loops of ijk.jl:108-136 (every contraction index spelled out);
 *                    O(o^3 v^4) scalar -- small shapes only.
 *   pt_oracle_gemm   the ijk2.jl organisation: per triplet six [v^2 x v].[v x v] + six
 *                    [v^2 x o].[o x v] GEMMs, permute-adds, V build, the identical a>=b>=c energy
 *                    loop; OpenMP over triplets (the reference threads over i, ijk.jl:49).  This is
 *                    the timed CPU baseline ("port") and the checker at larger shapes.  The GEMMs run
 *                    either on the register-blocked kernel below ("own") or, after pt_oracle_use_blas, on a
 *                    single-threaded cblas_dgemm of an OpenBLAS loaded at run time -- the reference's own
 *                    set-up (BLAS.set_num_threads(1) inside the threaded triplet loop, ijk.jl:45-46).
 *
 * PARITY PIN: Julia is not available in the build container, so the reference itself cannot be run.
 * The pin is: naive == gemm == numpy transcriptions (oracle/pt_numpy.py) == independent spin-orbital
 * brute force, plus the reference's printed water/STO-3G known answer reproduced through
 * oracle/mini_ccsd.py (examples/Juliacon2022.ipynb:613-615).
 *
 * All arrays are Julia column-major (first index fastest), Float64:
 *   T1[i,a] (o,v)  T2[i,j,a,b] (o,o,v,v)  OVVV[i,a,b,c]=(ia|bc) (o,v,v,v)
 *   OOOV[i,j,k,a]=(ij|ka) (o,o,o,v)  OVOV[i,a,j,b]=(ia|jb) (o,v,o,v)  fo (o)  fv (v)
 */
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long long i64;

#define T1_(i, a) T1[(i) + (i64)o * (a)]
#define T2_(i, j, a, b) T2[(i) + (i64)o * ((j) + (i64)o * ((a) + (i64)v * (b)))]
#define OVVV_(i, a, b, c) OVVV[(i) + (i64)o * ((a) + (i64)v * ((b) + (i64)v * (c)))]
#define OOOV_(i, j, k, a) OOOV[(i) + (i64)o * ((j) + (i64)o * ((k) + (i64)o * (a)))]
#define OVOV_(i, a, j, b) OVOV[(i) + (i64)o * ((a) + (i64)v * ((j) + (i64)o * (b)))]
#define W_(a, b, c) W[(a) + (i64)v * ((b) + (i64)v * (c))]
#define V_(a, b, c) V[(a) + (i64)v * ((b) + (i64)v * (c))]

/* ijk.jl:120-136 -- identical loop nest and arithmetic */
static double energy_loop(int v, const double *W, const double *V, const double *fv, double Dijk, int dij, int djk)
{
    double E = 0.0;
    for (int a = 0; a < v; a++) {
        double Dijka = Dijk - fv[a];
        for (int b = 0; b <= a; b++) {
            double Dijkab = Dijka - fv[b];
            int dab = (a == b);
            for (int c = 0; c <= b; c++) {
                double Dd = Dijkab - fv[c];
                int dbc = (b == c);
                double X = W_(a, b, c) * V_(a, b, c) + W_(a, c, b) * V_(a, c, b) + W_(b, a, c) * V_(b, a, c) +
                           W_(b, c, a) * V_(b, c, a) + W_(c, a, b) * V_(c, a, b) + W_(c, b, a) * V_(c, b, a);
                double Y = V_(a, b, c) + V_(b, c, a) + V_(c, a, b);
                double Z = V_(a, c, b) + V_(b, a, c) + V_(c, b, a);
                double Ef = (Y - 2 * Z) * (W_(a, b, c) + W_(b, c, a) + W_(c, a, b)) +
                            (Z - 2 * Y) * (W_(a, c, b) + W_(b, a, c) + W_(c, b, a)) + 3 * X;
                E += Ef * (2 - dij - djk) / (Dd * (1 + dab + dbc));
            }
        }
    }
    return E;
}

/* ---------------------------------------------------------------------------------------------
 * Literal transcription of ijk.jl:49-136 (permutes of :24-32 undone).
 * ------------------------------------------------------------------------------------------- */
int pt_oracle_naive(int o, int v, const double *T1, const double *T2, const double *OVVV, const double *OOOV,
                    const double *OVOV, const double *fo, const double *fv, double *Et)
{
    i64 v3 = (i64)v * v * v;
    double *W = (double *)malloc(sizeof(double) * v3), *V = (double *)malloc(sizeof(double) * v3);
    if (!W || !V) return 1;
    double E = 0.0;
    for (int i = 0; i < o; i++)
        for (int j = 0; j <= i; j++)
            for (int k = 0; k <= j; k++) {
                for (int c = 0; c < v; c++)
                    for (int b = 0; b < v; b++)
                        for (int a = 0; a < v; a++) {
                            double w = 0.0;
                            for (int d = 0; d < v; d++) {
                                w += OVVV_(i, a, b, d) * T2_(k, j, c, d); /* :109 */
                                w += OVVV_(j, b, a, d) * T2_(k, i, c, d); /* :110 */
                                w += OVVV_(k, c, a, d) * T2_(j, i, b, d); /* :111 */
                                w += OVVV_(k, c, b, d) * T2_(j, i, d, a); /* :112 */
                                w += OVVV_(i, a, c, d) * T2_(k, j, d, b); /* :113 */
                                w += OVVV_(j, b, c, d) * T2_(k, i, d, a); /* :114 */
                            }
                            for (int l = 0; l < o; l++) {
                                w -= OOOV_(l, i, j, b) * T2_(k, l, c, a); /* :109 */
                                w -= OOOV_(l, j, i, a) * T2_(k, l, c, b); /* :110 */
                                w -= OOOV_(l, k, i, a) * T2_(j, l, b, c); /* :111 */
                                w -= OOOV_(l, k, j, b) * T2_(i, l, a, c); /* :112 */
                                w -= OOOV_(l, i, k, c) * T2_(j, l, b, a); /* :113 */
                                w -= OOOV_(l, j, k, c) * T2_(i, l, a, b); /* :114 */
                            }
                            W_(a, b, c) = w;
                            V_(a, b, c) = w + T1_(i, a) * OVOV_(j, b, k, c) + OVOV_(i, a, k, c) * T1_(j, b) +
                                          OVOV_(i, a, j, b) * T1_(k, c); /* :116 */
                        }
                E += energy_loop(v, W, V, fv, fo[i] + fo[j] + fo[k], i == j, j == k);
            }
    free(W);
    free(V);
    *Et = E; /* :145 */
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Small column-major DGEMM:  C(MxN) = alpha * A(MxK) . B(KxN) + beta * C,  lda=M, ldb=K, ldc=M.
 * Plays the role of Octavian.matmul_serial! (ijk2.jl:111-150).  8x6 register block on GCC vector types.
 * ------------------------------------------------------------------------------------------- */
typedef double v4d __attribute__((vector_size(32), aligned(8)));

static inline void micro_8x6(i64 M, int K, const double *A, const double *B, int ldb, double *C, double alpha, double beta,
                             int nn)
{
    v4d c[6][2];
    for (int n = 0; n < 6; n++) c[n][0] = c[n][1] = (v4d){0, 0, 0, 0};
    for (int k = 0; k < K; k++) {
        v4d a0 = *(const v4d *)(A + (i64)k * M), a1 = *(const v4d *)(A + (i64)k * M + 4);
        for (int n = 0; n < 6; n++) {
            double b = (n < nn) ? B[k + (i64)ldb * n] : 0.0;
            v4d bb = {b, b, b, b};
            c[n][0] += a0 * bb;
            c[n][1] += a1 * bb;
        }
    }
    for (int n = 0; n < nn; n++) {
        v4d *c0 = (v4d *)(C + (i64)n * M), *c1 = (v4d *)(C + (i64)n * M + 4);
        v4d al = {alpha, alpha, alpha, alpha}, be = {beta, beta, beta, beta};
        if (beta == 0.0) {
            *c0 = al * c[n][0];
            *c1 = al * c[n][1];
        } else {
            *c0 = al * c[n][0] + be * *c0;
            *c1 = al * c[n][1] + be * *c1;
        }
    }
}

/* optional external BLAS (column-major cblas_dgemm, 32-bit integers), see pt_oracle_use_blas */
typedef void (*cblas_dgemm_fn)(int, int, int, int, int, int, double, const double *, int, const double *, int, double, double *, int);
static cblas_dgemm_fn ext_dgemm = NULL;

/* path == NULL or "": back to the built-in kernel.  Otherwise dlopen `path`, pin it to one thread per caller and route every
 * GEMM of pt_oracle_gemm through its cblas_dgemm.  Returns 0 on success. */
int pt_oracle_use_blas(const char *path)
{
    ext_dgemm = NULL;
    if (!path || !path[0]) return 0;
    void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) return 1;
    void (*set_threads)(int) = (void (*)(int))dlsym(lib, "scipy_openblas_set_num_threads");
    if (!set_threads) set_threads = (void (*)(int))dlsym(lib, "openblas_set_num_threads");
    cblas_dgemm_fn f = (cblas_dgemm_fn)dlsym(lib, "scipy_cblas_dgemm");
    if (!f) f = (cblas_dgemm_fn)dlsym(lib, "cblas_dgemm");
    if (!f) return 2;
    if (set_threads) set_threads(1);
    ext_dgemm = f;
    return 0;
}

static void dgemm_nn(i64 M, int N, int K, double alpha, const double *A, const double *B, double beta, double *C)
{
    if (ext_dgemm) { /* CblasColMajor = 102, CblasNoTrans = 111 */
        ext_dgemm(102, 111, 111, (int)M, N, K, alpha, A, (int)M, B, K, beta, C, (int)M);
        return;
    }
    const i64 MB = 96; /* rows per panel: 96*K*8 B stays in L2 */
    for (i64 m0 = 0; m0 < M; m0 += MB) {
        i64 m1 = m0 + MB < M ? m0 + MB : M;
        for (int n0 = 0; n0 < N; n0 += 6) {
            int nn = N - n0 < 6 ? N - n0 : 6;
            i64 m = m0;
            for (; m + 8 <= m1; m += 8) micro_8x6(M, K, A + m, B + (i64)K * n0, K, C + m + (i64)M * n0, alpha, beta, nn);
            for (; m < m1; m++) /* row tail */
                for (int n = 0; n < nn; n++) {
                    double s = 0.0;
                    for (int k = 0; k < K; k++) s += A[m + (i64)k * M] * B[k + (i64)K * (n0 + n)];
                    double *c = C + m + (i64)M * (n0 + n);
                    *c = (beta == 0.0) ? alpha * s : alpha * s + beta * *c;
                }
        }
    }
}

/* flattened triplet index t -> (i,j,k), i>=j>=k, k fastest (the order ijk.jl:49,63,83 visits them) */
static void triplet_from_index(i64 t, int *pi, int *pj, int *pk)
{
    int i = 0;
    while ((i64)(i + 1) * (i + 2) * (i + 3) / 6 <= t) i++;
    t -= (i64)i * (i + 1) * (i + 2) / 6;
    int j = 0;
    while ((i64)(j + 1) * (j + 2) / 2 <= t) j++;
    t -= (i64)j * (j + 1) / 2;
    *pi = i;
    *pj = j;
    *pk = (int)t;
}

/* ---------------------------------------------------------------------------------------------
 * ijk2.jl organisation.  Triplets [t_begin, t_end) of the flattened i>=j>=k list (t_end<0: all).
 * nthreads<=0: all OpenMP threads.  Returns the partial E(T) of that range.
 * ------------------------------------------------------------------------------------------- */
int pt_oracle_gemm(int o, int v, const double *T1, const double *T2, const double *OVVV, const double *OOOV,
                   const double *OVOV, const double *fo, const double *fv, i64 t_begin, i64 t_end, int nthreads,
                   double *Et)
{
    const i64 v2 = (i64)v * v, v3 = v2 * v;
    const i64 ntrip = (i64)o * (o + 1) * (o + 2) / 6;
    if (t_end < 0 || t_end > ntrip) t_end = ntrip;
    if (t_begin < 0) t_begin = 0;
    /* layout prep, ijk2.jl:26-33 */
    double *vvvo = (double *)malloc(sizeof(double) * v3 * o);     /* [(x,y),d,p] = OVVV[p,y,x,d] */
    double *T2p = (double *)malloc(sizeof(double) * v2 * o * o);  /* [d,c,q,r]   = T2[r,q,c,d]   */
    double *T2m = (double *)malloc(sizeof(double) * v2 * o * o);  /* [(x,y),l,p] = T2[p,l,y,x]   */
    double *ovoo = (double *)malloc(sizeof(double) * o * v * o * o); /* [l,c,q,r] = OOOV[l,q,r,c] */
    double *vvoo = (double *)malloc(sizeof(double) * v2 * o * o); /* [a,b,q,r]   = OVOV[q,a,r,b] */
    if (!vvvo || !T2p || !T2m || !ovoo || !vvoo) return 1;
#pragma omp parallel for collapse(2)
    for (int p = 0; p < o; p++)
        for (int d = 0; d < v; d++)
            for (int y = 0; y < v; y++)
                for (int x = 0; x < v; x++) vvvo[x + v * y + v2 * (d + (i64)v * p)] = OVVV_(p, y, x, d);
#pragma omp parallel for collapse(2)
    for (int r = 0; r < o; r++)
        for (int q = 0; q < o; q++)
            for (int c = 0; c < v; c++)
                for (int d = 0; d < v; d++) {
                    T2p[d + v * c + v2 * (q + (i64)o * r)] = T2_(r, q, c, d);
                    vvoo[d + v * c + v2 * (q + (i64)o * r)] = OVOV_(q, d, r, c);
                }
#pragma omp parallel for collapse(2)
    for (int p = 0; p < o; p++)
        for (int l = 0; l < o; l++)
            for (int y = 0; y < v; y++)
                for (int x = 0; x < v; x++) T2m[x + v * y + v2 * (l + (i64)o * p)] = T2_(p, l, y, x);
    for (int r = 0; r < o; r++)
        for (int q = 0; q < o; q++)
            for (int c = 0; c < v; c++)
                for (int l = 0; l < o; l++) ovoo[l + o * (c + (i64)v * (q + (i64)o * r))] = OOOV_(l, q, r, c);

    double Etot = 0.0;
    int fail = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : Etot)
    {
        double *W = (double *)malloc(sizeof(double) * v3), *V = (double *)malloc(sizeof(double) * v3),
               *X = (double *)malloc(sizeof(double) * v3);
        if (!W || !V || !X) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (i64 t = t_begin; t < t_end; t++) {
                int i, j, k;
                triplet_from_index(t, &i, &j, &k);
                if (i == j && j == k) continue; /* weight (2-dij-djk) = 0, ijk.jl:133 */
                const int P[6] = {j, k, i, k, i, j}, Q[6] = {i, i, j, j, k, k}, R[6] = {k, j, k, i, j, i};
                for (int term = 0; term < 6; term++) {
                    int p = P[term], q = Q[term], r = R[term];
                    /* X(p;q,r)[(x,y),z] = vvvo_p . T2p_qr - T2m_p . ovoo_qr    (ijk2.jl:111-112 etc.) */
                    dgemm_nn(v2, v, v, 1.0, vvvo + v3 * p, T2p + v2 * (q + (i64)o * r), 0.0, X);
                    dgemm_nn(v2, v, o, -1.0, T2m + v2 * o * (i64)p, ovoo + (i64)o * v * (q + (i64)o * r), 1.0, X);
#define X_(x, y, z) X[(x) + (i64)v * ((y) + (i64)v * (z))]
                    switch (term) { /* permutedims! + add, ijk2.jl:113-153 */
                    case 0: memcpy(W, X, sizeof(double) * v3); break;                                        /* abc */
                    case 1: for (int c = 0; c < v; c++) for (int b = 0; b < v; b++) for (int a = 0; a < v; a++) W_(a, b, c) += X_(a, c, b); break;
                    case 2: for (int c = 0; c < v; c++) for (int b = 0; b < v; b++) for (int a = 0; a < v; a++) W_(a, b, c) += X_(b, a, c); break;
                    case 3: for (int c = 0; c < v; c++) for (int b = 0; b < v; b++) for (int a = 0; a < v; a++) W_(a, b, c) += X_(b, c, a); break;
                    case 4: for (int c = 0; c < v; c++) for (int b = 0; b < v; b++) for (int a = 0; a < v; a++) W_(a, b, c) += X_(c, a, b); break;
                    case 5: for (int c = 0; c < v; c++) for (int b = 0; b < v; b++) for (int a = 0; a < v; a++) W_(a, b, c) += X_(c, b, a); break;
                    }
                }
                const double *vjk = vvoo + v2 * (j + (i64)o * k), *vik = vvoo + v2 * (i + (i64)o * k),
                             *vij = vvoo + v2 * (i + (i64)o * j);
                for (int c = 0; c < v; c++) /* ijk2.jl:155-157 */
                    for (int b = 0; b < v; b++)
                        for (int a = 0; a < v; a++)
                            V_(a, b, c) = W_(a, b, c) + T1_(i, a) * vjk[b + v * c] + vik[a + v * c] * T1_(j, b) +
                                          vij[a + v * b] * T1_(k, c);
                Etot += energy_loop(v, W, V, fv, fo[i] + fo[j] + fo[k], i == j, j == k);
            }
        }
        free(W);
        free(V);
        free(X);
    }
    free(vvvo);
    free(T2p);
    free(T2m);
    free(ovoo);
    free(vvoo);
    *Et = Etot;
    return fail;
}

int pt_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
