/* pt_oracle.c -- plain-C restatement of the reference's perturbative-triples loop.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline).  Nothing under sisi4s_b200/
 * links or calls this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * It follows, per sorted hole triple i<=j<=k,
 *   /root/reference/src/algorithms/CcsdPerturbativeTriples.cxx:159-216
 * with getDoublesContribution (:87-96, the vertex product replaced by the identical
 * PPPHCoulombIntegrals block, CoulombIntegralsFromVertex.cxx:430-431),
 * getSinglesContribution (:81-85), getEnergyDenominator (:98-117) and the
 * permutation algebra of src/math/Permutation.hpp:49-101.  The two contractions of
 * getDoublesContribution are done as blocked GEMMs (the reference delegates them to
 * Cyclops CTF @53ae5daa + BLAS, un-vendored), everything else is literal.
 *
 * PARITY: pinned against oracle/pt_oracle.py (NumPy forms A/B) in tests/test_oracle_c.py and,
 * like them, against the (T) energy the reference records for the UEG test system
 * (cc4s.correct.out.yaml:166-169) in tests/test_known_answers.py: agreement 3e-12.
 *
 * All arrays column-major in the reference's CTF index order:
 *   T1[a,i] T2[a,b,i,j] Vpphh[a,b,i,j] Vhhhp[i,j,k,a] Vppph[a,b,c,i].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Permutation<3>(p).images, Permutation.hpp:52-62 */
static const int PERM[6][3] = {{0, 1, 2}, {1, 0, 2}, {1, 2, 0}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}};
/* spinAndFermiFactors[invariantElementsCount(sigma_s)], :143,202 */
static const double SF[6] = {8.0, -4.0, 2.0, -4.0, 2.0, -4.0};

/* Optional BLAS: the reference hands these two contractions to Cyclops CTF, which calls the system
 * dgemm_.  oracle_set_dgemm() installs a Fortran-interface dgemm (the loader passes the one of the
 * OpenBLAS that ships with scipy, oracle/c_oracle.py:use_blas) so that the CPU baseline runs on a
 * vendor-grade GEMM; without it the blocked loop below is used. */
typedef void (*dgemm_fn)(char *, char *, int *, int *, int *, double *, double *, int *, double *, int *,
                         double *, double *, int *);
static dgemm_fn g_dgemm = 0;
void oracle_set_dgemm(void *fn) { g_dgemm = (dgemm_fn)fn; }
int oracle_has_dgemm(void) { return g_dgemm != 0; }

/* C[m + M*n] (+)= sum_k A[m + lda*k] * B[n + ldb*k]   (A, B "MN-major"; C column-major) */
static void gemm_nt(int M, int N, int K, const double *A, size_t lda, const double *B, size_t ldb,
                    double *C, double alpha, int accumulate) {
  if (g_dgemm && lda < 2147483647u && ldb < 2147483647u) {
    char tn = 'N', tt = 'T';
    int ilda = (int)lda, ildb = (int)ldb, ldc = M;
    double beta = accumulate ? 1.0 : 0.0;
    g_dgemm(&tn, &tt, &M, &N, &K, &alpha, (double *)A, &ilda, (double *)B, &ildb, &beta, C, &ldc);
    return;
  }
#pragma omp parallel for schedule(static)
  for (int n0 = 0; n0 < N; n0 += 4) {
    const int nb = N - n0 < 4 ? N - n0 : 4;
    double *c0 = C + (size_t)M * n0;
    if (!accumulate) memset(c0, 0, sizeof(double) * (size_t)M * nb);
    if (nb == 4) {
      /* 16 x 4 register block, K innermost */
      int m0 = 0;
      for (; m0 + 16 <= M; m0 += 16) {
        double acc[4][16];
        for (int jj = 0; jj < 4; ++jj)
          for (int mm = 0; mm < 16; ++mm) acc[jj][mm] = 0.0;
        for (int k = 0; k < K; ++k) {
          const double *a = A + lda * k + m0;
          const double *b = B + n0 + ldb * k;
          const double b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];
#pragma omp simd
          for (int mm = 0; mm < 16; ++mm) {
            const double av = a[mm];
            acc[0][mm] += av * b0;
            acc[1][mm] += av * b1;
            acc[2][mm] += av * b2;
            acc[3][mm] += av * b3;
          }
        }
        for (int jj = 0; jj < 4; ++jj)
          for (int mm = 0; mm < 16; ++mm) c0[(size_t)M * jj + m0 + mm] += alpha * acc[jj][mm];
      }
      if (m0 < M) {
        double *c1 = c0 + M, *c2 = c1 + M, *c3 = c2 + M;
        for (int k = 0; k < K; ++k) {
          const double *a = A + lda * k;
          const double b0 = alpha * B[n0 + ldb * k], b1 = alpha * B[n0 + 1 + ldb * k];
          const double b2 = alpha * B[n0 + 2 + ldb * k], b3 = alpha * B[n0 + 3 + ldb * k];
          for (int m = m0; m < M; ++m) {
            const double av = a[m];
            c0[m] += av * b0;
            c1[m] += av * b1;
            c2[m] += av * b2;
            c3[m] += av * b3;
          }
        }
      }
    } else {
      for (int j = 0; j < nb; ++j) {
        double *c = c0 + (size_t)M * j;
        for (int k = 0; k < K; ++k) {
          const double *a = A + lda * k;
          const double b = alpha * B[n0 + j + ldb * k];
          for (int m = 0; m < M; ++m) c[m] += a[m] * b;
        }
      }
    }
  }
}

/* Where the two big integral tensors live: either the dense reference tensors, or -- for shapes whose
 * v^3 o / v^2 o^2 tensors are not wanted on the host (o=64 v=512, o=100 v=800) -- tables of pointers
 * to the hole slabs Vppph[:,:,:,z] and the hole-pair blocks Vpphh[:,:,j,k] that the listed triples touch. */
typedef struct {
  const double *Vpphh, *Vppph;
  const double *const *pphh_blocks; /* [j + o*k] -> v^2 block, or NULL table */
  const double *const *ppph_slabs;  /* [z] -> v^3 slab, or NULL table */
} IntegralSource;
static const double *slab_of(const IntegralSource *src, int v, int z) {
  return src->ppph_slabs ? src->ppph_slabs[z] : src->Vppph + (size_t)v * v * v * z;
}
static const double *pair_block_of(const IntegralSource *src, int o, int v, int j, int k) {
  return src->pphh_blocks ? src->pphh_blocks[j + (size_t)o * k] : src->Vpphh + (size_t)v * v * (j + (size_t)o * k);
}

/* getDoublesContribution(x,y,z): W[a,b,c] = sum_d T2[a,d,x,y] V[b,c,d,z] - sum_l T2[a,b,x,l] Vhhhp[y,z,l,c] */
static void doubles_contribution(int o, int v, const double *T2, const IntegralSource *src, const double *Vhhhp,
                                 int x, int y, int z, double *W, double *upanel) {
  const size_t vv = (size_t)v * v;
  /* particle term: M=a, N=(b,c), K=d */
  gemm_nt(v, (int)vv, v, T2 + vv * ((size_t)x + (size_t)o * y), (size_t)v, slab_of(src, v, z), vv, W, 1.0, 0);
  /* hole term: M=(a,b), N=c, K=l;  A[(a,b) + v^2 o * l] = T2[a,b,x,l],  B[c + v*l] = Vhhhp[y,z,l,c] */
  for (int l = 0; l < o; ++l)
    for (int c = 0; c < v; ++c)
      upanel[c + (size_t)v * l] = Vhhhp[y + (size_t)o * (z + (size_t)o * (l + (size_t)o * c))];
  gemm_nt((int)vv, v, o, T2 + vv * x, vv * o, upanel, (size_t)v, W, -1.0, 1);
}

/* energy contribution of one sorted triple; scratch: 6 v^3 W blocks + v*o panel */
static double triple_energy(int o, int v, const double *epsi, const double *epsa, const double *T1,
                            const double *T2, const IntegralSource *src, const double *Vhhhp,
                            int i, int j, int k, double *scratch) {
  const size_t n3 = (size_t)v * v * v, vv = (size_t)v * v;
  const int h[3] = {i, j, k};
  int hp[6][3], distinct[6], rep[6];
  double *Wp[6];
  const double *Pp[6]; /* Vpphh[:,:,hp[p][1],hp[p][2]] */
  double *upanel = scratch + 6 * n3;
  for (int p = 0; p < 6; ++p) {
    for (int m = 0; m < 3; ++m) hp[p][m] = h[PERM[p][m]];
    int q = 0;
    for (; q < p; ++q)
      if (hp[q][0] == hp[p][0] && hp[q][1] == hp[p][1] && hp[q][2] == hp[p][2]) break;
    distinct[p] = (q == p);
    rep[p] = q;
    Wp[p] = scratch + n3 * q; /* duplicates reuse the earlier block (:170-173) */
    Pp[p] = distinct[p] ? pair_block_of(src, o, v, hp[p][1], hp[p][2]) : 0;
    if (distinct[p]) doubles_contribution(o, v, T2, src, Vhhhp, hp[p][0], hp[p][1], hp[p][2], Wp[p], upanel);
  }
  (void)rep;
  const double e3 = epsi[i] + epsi[j] + epsi[k];
  double e = 0.0;
#pragma omp parallel for reduction(+ : e) schedule(static)
  for (int c = 0; c < v; ++c) {
    for (int b = 0; b < v; ++b) {
      for (int a = 0; a < v; ++a) {
        const int xx[3] = {a, b, c};
        /* DVabc["abc"] += piDVabc[p]["abc" o pi]  (:179) */
        double X = 0.0;
        for (int p = 0; p < 6; ++p)
          X += Wp[p][xx[PERM[p][0]] + (size_t)v * xx[PERM[p][1]] + vv * xx[PERM[p][2]]];
        X = X / (e3 - epsa[a] - epsa[b] - epsa[c]); /* :183-191 */
        for (int p = 0; p < 6; ++p) {
          if (!distinct[p]) continue;
          double Y = 0.0;
          for (int s = 0; s < 6; ++s) {
            /* ("abc" o sigma) o pi : coordinate m is x[sigma(pi(m))]  (:205-211) */
            const int c0 = xx[PERM[s][PERM[p][0]]], c1 = xx[PERM[s][PERM[p][1]]], c2 = xx[PERM[s][PERM[p][2]]];
            const double wv = Wp[p][c0 + (size_t)v * c1 + vv * c2];
            const double sv = 0.5 * T1[c0 + (size_t)v * hp[p][0]] * Pp[p][c1 + (size_t)v * c2];
            Y += SF[s] * (wv + sv);
          }
          e += X * Y; /* :214 */
        }
      }
    }
  }
  return e;
}

/* index t of the reference enumeration (:156-158) -> (i,j,k) */
static void triple_of(int o, int64_t t, int *pi, int *pj, int *pk) {
  int64_t n = 0;
  for (int i = 0; i < o; ++i)
    for (int j = i; j < o; ++j) {
      const int64_t cnt = o - j;
      if (t < n + cnt) {
        *pi = i; *pj = j; *pk = j + (int)(t - n);
        return;
      }
      n += cnt;
    }
  *pi = *pj = *pk = o - 1;
}

static int triples_list(int o, int v, const double *epsi, const double *epsa, const double *T1, const double *T2,
                        const IntegralSource *src, const double *Vhhhp, const int64_t *idx, int64_t n,
                        double *e_per_triple, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  const size_t n3 = (size_t)v * v * v;
  double *scratch = (double *)malloc(sizeof(double) * (6 * n3 + (size_t)v * o));
  if (!scratch) return -1;
  for (int64_t q = 0; q < n; ++q) {
    int i, j, k;
    triple_of(o, idx[q], &i, &j, &k);
    e_per_triple[q] = triple_energy(o, v, epsi, epsa, T1, T2, src, Vhhhp, i, j, k, scratch);
  }
  free(scratch);
  return 0;
}

/* E_t for the sorted triples listed in idx[0..n); returns 0 or -1 (allocation failure) */
int oracle_triples_list(int o, int v, const double *epsi, const double *epsa, const double *T1,
                        const double *T2, const double *Vpphh, const double *Vhhhp, const double *Vppph,
                        const int64_t *idx, int64_t n, double *e_per_triple, int nthreads) {
  const IntegralSource src = {Vpphh, Vppph, 0, 0};
  return triples_list(o, v, epsi, epsa, T1, T2, &src, Vhhhp, idx, n, e_per_triple, nthreads);
}

/* the same with the PPPH slabs / PPHH hole-pair blocks given as pointer tables (entries the listed
 * triples do not touch may be NULL) */
int oracle_triples_list_blocks(int o, int v, const double *epsi, const double *epsa, const double *T1,
                               const double *T2, const double *const *pphh_blocks, const double *Vhhhp,
                               const double *const *ppph_slabs, const int64_t *idx, int64_t n,
                               double *e_per_triple, int nthreads) {
  const IntegralSource src = {0, 0, pphh_blocks, ppph_slabs};
  return triples_list(o, v, epsi, epsa, T1, T2, &src, Vhhhp, idx, n, e_per_triple, nthreads);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
