/*
 * spike_oracle.c -- CPU ORACLE (plain C, fp64).  TEST INFRASTRUCTURE ONLY; see spike_oracle.h.
 *
 * Every function that restates reference code cites the reference file:line it follows
 * (paths relative to /root/reference).  Nothing here is copied from the reference; the logic is
 * re-expressed over plain CSR / band arrays because PETSc is not available in this image.
 */
#include "spike_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BW(k) (2 * (int64_t)(k) + 1)
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* =========================================================================================
 * Synthetic inputs: counter-based, identical bits on CPU and GPU.
 * ========================================================================================= */
uint64_t orc_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
double orc_u01(uint64_t seed, uint64_t counter) {
  return (double)(orc_splitmix64(seed ^ counter) >> 11) * (1.0 / 9007199254740992.0);
}
/* a_ij = 2 u01(seed ^ (i(2k+1)+d+k)) - 1 for d != 0; a_ii = delta * sum_{d!=0}|a_ij| (d ascending). */
void orc_gen_band(int64_t n, int k, uint64_t seed, double delta, double *a) {
  const int64_t bw = BW(k);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    double *row = a + i * bw;
    for (int d = -k; d <= k; ++d) {
      const int64_t j = i + d;
      if (d == 0) continue;
      if (j < 0 || j >= n) { row[d + k] = 0.0; continue; }
      const double v = 2.0 * orc_u01(seed, (uint64_t)(i * bw + d + k)) - 1.0;
      row[d + k] = v;
      s += fabs(v);
    }
    row[k] = delta * s;
    if (row[k] == 0.0) row[k] = 1.0;
  }
}
void orc_gen_vec(int64_t n, uint64_t seed, double *u) {
  for (int64_t i = 0; i < n; ++i) u[i] = orc_u01(seed ^ 0xA5A5A5A5A5A5A5A5ULL, (uint64_t)i);
}

/* =========================================================================================
 * MatCreateSubMatrixBanded -- /root/reference/src/matbanded.c:22-107
 * ========================================================================================= */
/* (i) weights in row-major traversal order, matbanded.c:38-49; (ii) k search, :53-56;
 * returned values, :104-105.  Keeps the fall-through quirk: if the loop never breaks k == kmax
 * and normB excludes w[kmax].  The reference has no bound check kmax <= n (hazard noted in
 * SURVEY 8a-2); the oracle returns -1 at the point where the reference would read w[n]. */
int orc_band_select(int n, const int *ia, const int *ja, const double *a, int kmax, double frac,
                    int *k_out, double *frac_out) {
  double *w = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double normA = 0.0, normB = 0.0;
  for (int r = 0; r < n; ++r)
    for (int c = ia[r]; c < ia[r + 1]; ++c) {
      w[abs(r - ja[c])] += fabs(a[c]);
      normA += fabs(a[c]);
    }
  int k;
  for (k = 0; k < kmax; ++k) {
    if (k >= n) { free(w); return -1; }
    normB += w[k];
    if (normB >= frac * normA) break;
  }
  free(w);
  *k_out = k;
  *frac_out = normB / normA;
  return 0;
}
/* copy loop, matbanded.c:84-99 (column order preserved) */
int64_t orc_band_extract_csr(int n, const int *ia, const int *ja, const double *a, int k, int *ib,
                             int *jb, double *b) {
  int64_t nnz = 0;
  ib[0] = 0;
  for (int r = 0; r < n; ++r) {
    for (int c = ia[r]; c < ia[r + 1]; ++c) {
      if (abs(ja[c] - r) > k) continue;
      jb[nnz] = ja[c];
      b[nnz] = a[c];
      ++nnz;
    }
    ib[r + 1] = (int)nnz;
  }
  return nnz;
}
void orc_csr_to_band(int n, const int *ia, const int *ja, const double *a, int k, double *band) {
  const int64_t bw = BW(k);
  memset(band, 0, sizeof(double) * (size_t)n * (size_t)bw);
  for (int r = 0; r < n; ++r)
    for (int c = ia[r]; c < ia[r + 1]; ++c) {
      const int d = ja[c] - r;
      if (abs(d) > k) continue;
      band[(int64_t)r * bw + d + k] += a[c];
    }
}

/* =========================================================================================
 * MatPermute / VecPermute semantics at /root/reference/src/kspreorder.c:20,122-127
 * [EXTERNAL: PETSc] B(i,j) = A(rowp[i], colp[j]); new[i] = old[idx[i]] (inverse: new[idx[i]] = old[i])
 * ========================================================================================= */
typedef struct { int c; double v; } colval;
static int cmp_colval(const void *x, const void *y) {
  const int a = ((const colval *)x)->c, b = ((const colval *)y)->c;
  return (a > b) - (a < b);
}
int orc_mat_permute_csr(int n, const int *ia, const int *ja, const double *a, const int *rowp,
                        const int *colp, int *ib, int *jb, double *b) {
  int *icol = (int *)malloc(sizeof(int) * (size_t)n);
  for (int j = 0; j < n; ++j) icol[j] = -1;
  for (int j = 0; j < n; ++j) {
    if (colp[j] < 0 || colp[j] >= n || icol[colp[j]] >= 0) { free(icol); return -1; }
    icol[colp[j]] = j;
  }
  int maxrow = 0;
  for (int r = 0; r < n; ++r) if (ia[r + 1] - ia[r] > maxrow) maxrow = ia[r + 1] - ia[r];
  colval *tmp = (colval *)malloc(sizeof(colval) * (size_t)(maxrow > 0 ? maxrow : 1));
  int64_t nnz = 0;
  ib[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int r = rowp[i];
    if (r < 0 || r >= n) { free(icol); free(tmp); return -1; }
    const int len = ia[r + 1] - ia[r];
    for (int c = 0; c < len; ++c) { tmp[c].c = icol[ja[ia[r] + c]]; tmp[c].v = a[ia[r] + c]; }
    qsort(tmp, (size_t)len, sizeof(colval), cmp_colval);
    for (int c = 0; c < len; ++c) { jb[nnz] = tmp[c].c; b[nnz] = tmp[c].v; ++nnz; }
    ib[i + 1] = (int)nnz;
  }
  free(icol); free(tmp);
  return 0;
}
void orc_vec_permute(int n, double *x, const int *idx, int inverse) {
  double *t = (double *)malloc(sizeof(double) * (size_t)n);
  if (!inverse) for (int i = 0; i < n; ++i) t[i] = x[idx[i]];
  else          for (int i = 0; i < n; ++i) t[idx[i]] = x[i];
  memcpy(x, t, sizeof(double) * (size_t)n);
  free(t);
}

/* =========================================================================================
 * MatGetOrdering_AWBM -- /root/reference/src/petsc_mat_awbm.c:42-225
 * The reference treats the CSR arrays "as if the matrix were column-major" (:47): index c walks
 * CSR rows, ja[] entries are called rows.  Restated phase by phase with the same scan orders.
 * ========================================================================================= */
int orc_awbm(int n, const int *ia, const int *ja, const double *a, int *permR, int *match_out) {
  const double eps = sqrt(DBL_EPSILON); /* PETSC_SQRT_MACHINE_EPSILON, :59 */
  const int64_t nnz = ia[n];
  int *match = (int *)malloc(sizeof(int) * (size_t)n), *matchR = (int *)malloc(sizeof(int) * (size_t)n);
  double *amax = (double *)calloc((size_t)n, sizeof(double));
  double *u = (double *)calloc((size_t)n, sizeof(double)), *v = (double *)calloc((size_t)n, sizeof(double));
  double *w = (double *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(double));
  int c, r, r1, c1, rc = 0;
  /* MatGetRowMaxAbs, :66 */
  for (c = 0; c < n; ++c) for (r = ia[c]; r < ia[c + 1]; ++r) if (fabs(a[r]) > amax[c]) amax[c] = fabs(a[r]);
  for (c = 0; c < n; ++c) match[c] = -1;
  /* weights, :73-80 */
  for (c = 0; c < n; ++c)
    for (r = ia[c]; r < ia[c + 1]; ++r) {
      const double ar = fabs(a[r]);
      w[r] = (ar == 0.0) ? DBL_MAX : log(amax[c] / ar);
    }
  /* row duals, :82-87 */
  for (r = 0; r < n; ++r) u[r] = DBL_MAX;
  for (c = 0; c < n; ++c) for (r = ia[c]; r < ia[c + 1]; ++r) if (w[r] < u[ja[r]]) u[ja[r]] = w[r];
  /* column duals, :89-94 */
  for (c = 0; c < n; ++c) {
    v[c] = DBL_MAX;
    for (r = ia[c]; r < ia[c + 1]; ++r) if (w[r] - u[ja[r]] < v[c]) v[c] = w[r] - u[ja[r]];
  }
  for (r = 0; r < n; ++r) matchR[r] = -1;
  /* greedy tight-edge matching, :98-112 */
  for (c = 0; c < n; ++c)
    for (r = ia[c]; r < ia[c + 1]; ++r) {
      const double wt = w[r] - u[ja[r]] - v[c];
      if (wt <= eps && matchR[ja[r]] < 0) { match[c] = ja[r]; matchR[ja[r]] = c; break; }
    }
  /* one-level augmentation over tight edges, :115-140 */
  for (c = 0; c < n; ++c) {
    if (match[c] >= 0) continue;
    for (r = ia[c]; r < ia[c + 1]; ++r) {
      const double wt = w[r] - u[ja[r]] - v[c];
      if (wt > eps) continue;
      c1 = matchR[ja[r]];
      /* the reference indexes ia[c1] without checking c1 >= 0 (:121-122); a tight edge to an
       * unmatched row cannot survive the greedy phase for this c, so c1 >= 0 here. */
      if (c1 < 0) continue;
      for (r1 = ia[c1]; r1 < ia[c1 + 1]; ++r1) {
        const double wt1 = w[r1] - u[ja[r1]] - v[c1];
        if (matchR[ja[r1]] < 0 && wt1 <= eps) {
          match[c] = ja[r]; matchR[ja[r]] = c; match[c1] = ja[r1]; matchR[ja[r1]] = c1;
          break;
        }
      }
      if (match[c] >= 0) break;
    }
  }
  /* non-optimal rows, :143-153 */
  for (c = 0; c < n; ++c) {
    if (match[c] >= 0) continue;
    for (r = ia[c]; r < ia[c + 1]; ++r)
      if (matchR[ja[r]] < 0) { match[c] = ja[r]; matchR[ja[r]] = c; break; }
  }
  /* non-optimal one-level augmentation, :156-178 */
  for (c = 0; c < n; ++c) {
    if (match[c] >= 0) continue;
    for (r = ia[c]; r < ia[c + 1]; ++r) {
      c1 = matchR[ja[r]];
      if (c1 < 0) continue;
      for (r1 = ia[c1]; r1 < ia[c1 + 1]; ++r1)
        if (matchR[ja[r1]] < 0) {
          match[c] = ja[r]; matchR[ja[r]] = c; match[c1] = ja[r1]; matchR[ja[r1]] = c1;
          break;
        }
      if (match[c] >= 0) break;
    }
  }
  /* completion, :181-193 (r persists across columns) */
  for (c = 0, r = 0; c < n; ++c) {
    if (match[c] >= n) { rc = -2; goto done; }
    if (match[c] < 0)
      for (; r < n; ++r)
        if (matchR[r] < 0) { match[c] = r; matchR[r] = c; break; }
  }
  /* check, :196-199 */
  for (c = 0; c < n; ++c) if (match[c] < 0 || match[c] >= n) { rc = -3; goto done; }
  /* permutation, :201 */
  for (c = 0; c < n; ++c) permR[match[c]] = c;
  if (match_out) memcpy(match_out, match, sizeof(int) * (size_t)n);
done:
  free(match); free(matchR); free(amax); free(u); free(v); free(w);
  return rc;
}

/* =========================================================================================
 * Exact banded solve = what "-banded_pc_type lu" with natural ordering computes on the extracted
 * band (call sites /root/reference/src/matbanded.c:178 [factor] and :190 [apply]).
 * [EXTERNAL arithmetic: PETSc LU is not vendored; restated as textbook no-pivot band LU with the
 *  SpikeGPU-style diagonal boosting named by north_star.]   Rows layout, in place.
 * ========================================================================================= */
static int64_t band_lu_range(int64_t lo, int64_t hi, int k, double *a, double boost) {
  /* factor the diagonal block rows/cols [lo,hi) of the band, ignoring couplings outside it */
  const int64_t bw = BW(k);
  int64_t nboost = 0;
  for (int64_t j = lo; j < hi; ++j) {
    double *rj = a + j * bw;
    double piv = rj[k];
    if (fabs(piv) < boost) { piv = (piv < 0.0) ? -boost : boost; rj[k] = piv; ++nboost; }
    const int64_t iend = imin64(j + k, hi - 1);
    const int64_t cend = imin64(j + k, hi - 1);
    for (int64_t i = j + 1; i <= iend; ++i) {
      double *ri = a + i * bw;
      const double l = ri[j - i + k] / piv;
      ri[j - i + k] = l;
      for (int64_t c = j + 1; c <= cend; ++c) ri[c - i + k] -= l * rj[c - j + k];
    }
  }
  return nboost;
}
/* solve with the factors of block [lo,hi); x is the full-length vector (entries lo..hi-1 used) */
static void band_solve_range(int64_t lo, int64_t hi, int k, const double *lu, double *x) {
  const int64_t bw = BW(k);
  for (int64_t i = lo; i < hi; ++i) {
    const double *ri = lu + i * bw;
    double s = x[i];
    for (int64_t j = imax64(lo, i - k); j < i; ++j) s -= ri[j - i + k] * x[j];
    x[i] = s;
  }
  for (int64_t i = hi - 1; i >= lo; --i) {
    const double *ri = lu + i * bw;
    double s = x[i];
    for (int64_t j = i + 1; j <= imin64(hi - 1, i + k); ++j) s -= ri[j - i + k] * x[j];
    x[i] = s / ri[k];
  }
}
int64_t orc_band_lu(int64_t n, int k, double *a, double boost) { return band_lu_range(0, n, k, a, boost); }
void orc_band_solve(int64_t n, int k, const double *lu, double *x, int nrhs, int64_t ldx) {
  for (int r = 0; r < nrhs; ++r) band_solve_range(0, n, k, lu, x + (int64_t)r * ldx);
}
void orc_band_mult(int64_t n, int k, const double *a, const double *x, double *y) {
  const int64_t bw = BW(k);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double *ri = a + i * bw;
    double s = 0.0;
    for (int64_t j = imax64(0, i - k); j <= imin64(n - 1, i + k); ++j) s += ri[j - i + k] * x[j];
    y[i] = s;
  }
}

/* =========================================================================================
 * Truncated SPIKE on the CPU (partition-parallel).  [EXTERNAL: Polizzi & Sameh; SaP::GPU --
 * the reference names SPIKE in README.md:4 but ships no code, SURVEY.md 8c.]
 *   A = D S,  D = diag(A_i).  V_i^(b) = bottom kxk of A_i^{-1}[0;B_i],  W_i^(t) = top kxk of
 *   A_i^{-1}[C_i;0].  Reduced blocks [I V_i^(b); W_{i+1}^(t) I].  Same algorithm as the GPU path.
 * ========================================================================================= */
struct orc_spike {
  int64_t n; int k; int nparts; int64_t tip_rows; double boost;
  int64_t *start;   /* nparts+1 */
  double *vb, *wt;  /* (nparts-1) * k*k each, row-major; interface i couples part i and i+1 */
  double *red;      /* (nparts-1) * k*k : LU (partial pivoting) of I - W V */
  int *piv;         /* (nparts-1) * k */
};

orc_spike *orc_spike_create(int64_t n, int k, int nparts, int align, int64_t tip_rows, double boost) {
  orc_spike *s = (orc_spike *)calloc(1, sizeof(orc_spike));
  s->n = n; s->k = k; s->nparts = nparts; s->tip_rows = tip_rows; s->boost = boost;
  s->start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nparts + 1));
  const int64_t units = (n + align - 1) / align;
  for (int p = 0; p <= nparts; ++p) s->start[p] = imin64(n, (units * p / nparts) * align);
  s->start[nparts] = n;
  const size_t kk = (size_t)k * (size_t)k, ni = (size_t)(nparts > 1 ? nparts - 1 : 1);
  s->vb = (double *)calloc(ni * kk, sizeof(double));
  s->wt = (double *)calloc(ni * kk, sizeof(double));
  s->red = (double *)calloc(ni * kk, sizeof(double));
  s->piv = (int *)calloc(ni * (size_t)k, sizeof(int));
  return s;
}
void orc_spike_destroy(orc_spike *s) {
  if (!s) return;
  free(s->start); free(s->vb); free(s->wt); free(s->red); free(s->piv); free(s);
}
const double *orc_spike_vb(const orc_spike *s, int i) { return s->vb + (size_t)i * s->k * s->k; }
const double *orc_spike_wt(const orc_spike *s, int i) { return s->wt + (size_t)i * s->k * s->k; }
int64_t orc_spike_part_start(const orc_spike *s, int p) { return s->start[p]; }

/* dense helpers (row-major kxk) */
static void dense_lu_piv(int k, double *m, int *piv) {
  for (int j = 0; j < k; ++j) {
    int p = j; double best = fabs(m[j * k + j]);
    for (int i = j + 1; i < k; ++i) if (fabs(m[i * k + j]) > best) { best = fabs(m[i * k + j]); p = i; }
    piv[j] = p;
    if (p != j) for (int c = 0; c < k; ++c) { double t = m[j * k + c]; m[j * k + c] = m[p * k + c]; m[p * k + c] = t; }
    const double d = m[j * k + j];
    for (int i = j + 1; i < k; ++i) {
      const double l = m[i * k + j] / d;
      m[i * k + j] = l;
      for (int c = j + 1; c < k; ++c) m[i * k + c] -= l * m[j * k + c];
    }
  }
}
static void dense_lu_solve(int k, const double *m, const int *piv, double *x) {
  for (int j = 0; j < k; ++j) { if (piv[j] != j) { double t = x[j]; x[j] = x[piv[j]]; x[piv[j]] = t; } }
  for (int i = 0; i < k; ++i) { double s = x[i]; for (int j = 0; j < i; ++j) s -= m[i * k + j] * x[j]; x[i] = s; }
  for (int i = k - 1; i >= 0; --i) { double s = x[i]; for (int j = i + 1; j < k; ++j) s -= m[i * k + j] * x[j]; x[i] = s / m[i * k + i]; }
}

/* UL elimination (bottom-up) of the leading m rows of partition [lo, lo+m) on a private copy,
 * down to the top kxk Schur block St = ((A_i[0:m,0:m])^{-1}_tt)^{-1}; then Wt = St^{-1} C_i. */
static void spike_wt_tip(const orc_spike *s, const double *a, int part, double *wt) {
  const int k = s->k; const int64_t bw = BW(k);
  const int64_t lo = s->start[part], hi_full = s->start[part + 1];
  int64_t m = hi_full - lo;
  if (s->tip_rows > 0 && s->tip_rows < m) m = s->tip_rows;
  if (m < k) m = imin64(hi_full - lo, k);
  const int64_t hi = lo + m;
  double *w = (double *)malloc(sizeof(double) * (size_t)m * (size_t)bw);
  memcpy(w, a + lo * bw, sizeof(double) * (size_t)m * (size_t)bw);
#define WR(i) (w + ((i) - lo) * bw)
  for (int64_t j = hi - 1; j >= lo + k; --j) {
    double *rj = WR(j);
    double piv = rj[k];
    if (fabs(piv) < s->boost) piv = (piv < 0.0) ? -s->boost : s->boost;
    const int64_t ibeg = imax64(lo, j - k);
    for (int64_t i = ibeg; i < j; ++i) {
      double *ri = WR(i);
      const double u = ri[j - i + k] / piv;
      for (int64_t c = ibeg; c < j; ++c) ri[c - i + k] -= u * rj[c - j + k];
    }
  }
  /* dense solve St * Wt = C_i, C_i(r,c) = A(lo+r, lo-k+c) (upper triangular incl. diagonal) */
  double *st = (double *)calloc((size_t)k * k, sizeof(double));
  int *piv = (int *)malloc(sizeof(int) * (size_t)k);
  for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) st[r * k + c] = WR(lo + r)[(lo + c) - (lo + r) + k];
#undef WR
  dense_lu_piv(k, st, piv);
  double *col = (double *)malloc(sizeof(double) * (size_t)k);
  for (int c = 0; c < k; ++c) {
    for (int r = 0; r < k; ++r) {
      const int64_t j = lo - k + c, i = lo + r;
      col[r] = (j >= 0 && j - i >= -k) ? a[i * bw + (j - i + k)] : 0.0;
    }
    dense_lu_solve(k, st, piv, col);
    for (int r = 0; r < k; ++r) wt[r * k + c] = col[r];
  }
  free(st); free(piv); free(col); free(w);
}

/* Vb = bottom kxk of A_i^{-1}[0;B_i] from the LU factors: (L_bb U_bb) Vb = B_i */
static void spike_vb_tip(const orc_spike *s, const double *lu, int part, double *vb) {
  const int k = s->k; const int64_t bw = BW(k);
  const int64_t hi = s->start[part + 1], lo = imax64(s->start[part], hi - k);
  double *buf = (double *)calloc((size_t)k + 1, sizeof(double));
  double *x = buf - lo; /* x[i] valid for global i in [lo,hi) */
  for (int c = 0; c < k; ++c) {
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t j = hi + c;
      x[i] = (j < s->n && j - i <= k) ? lu[i * bw + (j - i + k)] : 0.0;
    }
    band_solve_range(lo, hi, k, lu, x); /* factors of the trailing block equal the trailing factors */
    for (int r = 0; r < k; ++r) vb[r * k + c] = (hi - k + r >= lo) ? x[hi - k + r] : 0.0;
  }
  free(buf);
}

int64_t orc_spike_factor(orc_spike *s, double *a, int nthreads) {
  const int k = s->k; const int P = s->nparts;
  int64_t nboost = 0;
  (void)nthreads;
  /* W^(t) needs the unfactored rows: compute all top tips first (read-only on a), then LU in place */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int p = 1; p < P; ++p) spike_wt_tip(s, a, p, s->wt + (size_t)(p - 1) * k * k);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : nboost)
  for (int p = 0; p < P; ++p) nboost += band_lu_range(s->start[p], s->start[p + 1], k, a, s->boost);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int p = 0; p < P - 1; ++p) {
    double *vb = s->vb + (size_t)p * k * k, *wt = s->wt + (size_t)p * k * k;
    double *red = s->red + (size_t)p * k * k;
    spike_vb_tip(s, a, p, vb);
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) {
        double t = (r == c) ? 1.0 : 0.0;
        for (int q = 0; q < k; ++q) t -= wt[r * k + q] * vb[q * k + c];
        red[r * k + c] = t;
      }
    dense_lu_piv(k, red, s->piv + (size_t)p * k);
  }
  return nboost;
}

void orc_spike_solve(orc_spike *s, const double *lu, const double *b, double *x, int nthreads) {
  const int k = s->k; const int P = s->nparts; const int64_t n = s->n, bw = BW(k);
  (void)nthreads;
  if (x != b) memcpy(x, b, sizeof(double) * (size_t)n);
  /* (1) g = D^{-1} b */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int p = 0; p < P; ++p) band_solve_range(s->start[p], s->start[p + 1], k, lu, x);
  if (P == 1) return;
  /* (2) reduced system per interface: (I - W V) xt = gt - W gb ; xb = gb - V xt */
  double *xt = (double *)malloc(sizeof(double) * (size_t)(P - 1) * k);
  double *xb = (double *)malloc(sizeof(double) * (size_t)(P - 1) * k);
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int p = 0; p < P - 1; ++p) {
    const double *vb = s->vb + (size_t)p * k * k, *wt = s->wt + (size_t)p * k * k;
    const double *gb = x + s->start[p + 1] - k, *gt = x + s->start[p + 1];
    double *t = xt + (size_t)p * k, *bb = xb + (size_t)p * k;
    for (int r = 0; r < k; ++r) { double v = gt[r]; for (int c = 0; c < k; ++c) v -= wt[r * k + c] * gb[c]; t[r] = v; }
    dense_lu_solve(k, s->red + (size_t)p * k * k, s->piv + (size_t)p * k, t);
    for (int r = 0; r < k; ++r) { double v = gb[r]; for (int c = 0; c < k; ++c) v -= vb[r * k + c] * t[c]; bb[r] = v; }
  }
  /* (3) x_i = g_i - A_i^{-1}[C_i xb_{i-1}; 0] - A_i^{-1}[0; B_i xt_{i+1}], corrections restricted to
   *     tip_rows rows when truncation is on (spikes have decayed beyond that) */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int p = 0; p < P; ++p) {
    const int64_t lo = s->start[p], hi = s->start[p + 1], len = hi - lo;
    const int64_t m = (s->tip_rows > 0 && s->tip_rows < len) ? s->tip_rows : len;
    double *w = (double *)calloc((size_t)(len > 0 ? len : 1), sizeof(double));
    double *wv = w - lo; /* wv[i] for global i */
    if (p > 0) { /* top: rows lo..lo+k-1 get C_i xb_{p-1} */
      const double *bb = xb + (size_t)(p - 1) * k;
      for (int64_t i = lo; i < imin64(lo + k, hi); ++i) {
        double v = 0.0;
        for (int64_t j = imax64(0, i - k); j < lo; ++j) v += lu[i * bw + (j - i + k)] * bb[j - (lo - k)];
        wv[i] = v;
      }
      band_solve_range(lo, lo + m, k, lu, wv);
      for (int64_t i = lo; i < lo + m; ++i) { x[i] -= wv[i]; wv[i] = 0.0; }
    }
    if (p < P - 1) { /* bottom: rows hi-k..hi-1 get B_i xt_p */
      const double *t = xt + (size_t)p * k;
      for (int64_t i = imax64(lo, hi - k); i < hi; ++i) {
        double v = 0.0;
        for (int64_t j = hi; j <= imin64(n - 1, i + k); ++j) v += lu[i * bw + (j - i + k)] * t[j - hi];
        wv[i] = v;
      }
      band_solve_range(hi - m, hi, k, lu, wv);
      for (int64_t i = hi - m; i < hi; ++i) x[i] -= wv[i];
    }
    free(w);
  }
  free(xt); free(xb);
}

/* =========================================================================================
 * Krylov drivers standing in for the inner KSP at /root/reference/src/kspreorder.c:124.
 * [EXTERNAL: PETSc defaults] left preconditioning, preconditioned residual norm, x0 = 0,
 * GMRES restart 30 with classical Gram-Schmidt; convergence ||r_k|| <= rtol * ||M^{-1} b||.
 * ========================================================================================= */
static double vdot(int64_t n, const double *x, const double *y) { double s = 0; for (int64_t i = 0; i < n; ++i) s += x[i] * y[i]; return s; }
static double vnorm(int64_t n, const double *x) { return sqrt(vdot(n, x, x)); }

int orc_gmres(int64_t n, orc_apply_fn amul, void *actx, orc_apply_fn pc, void *pctx, const double *b,
              double *x, int restart, double rtol, int maxit, int *its, double *rnorm) {
  const int m = restart;
  double *V = (double *)malloc(sizeof(double) * (size_t)n * (size_t)(m + 1));
  double *H = (double *)calloc((size_t)(m + 1) * (size_t)m, sizeof(double));
  double *cs = (double *)calloc((size_t)m, sizeof(double)), *sn = (double *)calloc((size_t)m, sizeof(double));
  double *g = (double *)calloc((size_t)m + 1, sizeof(double)), *y = (double *)calloc((size_t)m, sizeof(double));
  double *t = (double *)malloc(sizeof(double) * (size_t)n), *z = (double *)malloc(sizeof(double) * (size_t)n);
  int it = 0, conv = 0;
  memset(x, 0, sizeof(double) * (size_t)n);
  pc(pctx, b, z);
  const double bnorm = vnorm(n, z);
  double res = bnorm;
  if (bnorm == 0.0) { conv = 1; goto done; }
  while (it < maxit && !conv) {
    /* r = M^{-1}(b - A x) */
    amul(actx, x, t);
    for (int64_t i = 0; i < n; ++i) t[i] = b[i] - t[i];
    pc(pctx, t, V);
    res = vnorm(n, V);
    if (res <= rtol * bnorm) { conv = 1; break; }
    for (int64_t i = 0; i < n; ++i) V[i] /= res;
    memset(g, 0, sizeof(double) * (size_t)(m + 1));
    g[0] = res;
    int j;
    for (j = 0; j < m && it < maxit; ++j) {
      double *vj1 = V + (size_t)(j + 1) * n;
      amul(actx, V + (size_t)j * n, t);
      pc(pctx, t, vj1);
      for (int i = 0; i <= j; ++i) H[i * m + j] = vdot(n, vj1, V + (size_t)i * n);        /* classical GS */
      for (int i = 0; i <= j; ++i) { const double h = H[i * m + j]; const double *vi = V + (size_t)i * n; for (int64_t q = 0; q < n; ++q) vj1[q] -= h * vi[q]; }
      const double hn = vnorm(n, vj1);
      H[(j + 1) * m + j] = hn;
      if (hn != 0.0) for (int64_t q = 0; q < n; ++q) vj1[q] /= hn;
      for (int i = 0; i < j; ++i) {
        const double a0 = H[i * m + j], a1 = H[(i + 1) * m + j];
        H[i * m + j] = cs[i] * a0 + sn[i] * a1;
        H[(i + 1) * m + j] = -sn[i] * a0 + cs[i] * a1;
      }
      const double a0 = H[j * m + j], a1 = H[(j + 1) * m + j], d = hypot(a0, a1);
      cs[j] = a0 / d; sn[j] = a1 / d;
      H[j * m + j] = d; H[(j + 1) * m + j] = 0.0;
      g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j];
      ++it;
      res = fabs(g[j + 1]);
      if (res <= rtol * bnorm) { conv = 1; ++j; break; }
    }
    const int jj = j;
    for (int i = jj - 1; i >= 0; --i) {
      double sacc = g[i];
      for (int c = i + 1; c < jj; ++c) sacc -= H[i * m + c] * y[c];
      y[i] = sacc / H[i * m + i];
    }
    for (int i = 0; i < jj; ++i) { const double *vi = V + (size_t)i * n; for (int64_t q = 0; q < n; ++q) x[q] += y[i] * vi[q]; }
  }
done:
  *its = it; *rnorm = res;
  free(V); free(H); free(cs); free(sn); free(g); free(y); free(t); free(z);
  return conv ? 0 : 1;
}

/* BiCGStab on the left-preconditioned system M^{-1}A x = M^{-1}b (PETSc KSPBCGS with PC_LEFT):
 * iteration k tests ||r_k|| (preconditioned) after the full step; 2 PC applies per iteration. */
int orc_bicgstab(int64_t n, orc_apply_fn amul, void *actx, orc_apply_fn pc, void *pctx,
                 const double *b, double *x, double rtol, int maxit, int *its, double *rnorm) {
  double *r = (double *)malloc(sizeof(double) * (size_t)n), *rh = (double *)malloc(sizeof(double) * (size_t)n);
  double *p = (double *)calloc((size_t)n, sizeof(double)), *v = (double *)calloc((size_t)n, sizeof(double));
  double *sv = (double *)malloc(sizeof(double) * (size_t)n), *t = (double *)malloc(sizeof(double) * (size_t)n);
  double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
  int it = 0, conv = 0;
  memset(x, 0, sizeof(double) * (size_t)n);
  pc(pctx, b, r);
  const double bnorm = vnorm(n, r);
  double res = bnorm, rho = 1.0, alpha = 1.0, omega = 1.0;
  memcpy(rh, r, sizeof(double) * (size_t)n);
  if (bnorm == 0.0) conv = 1;
  while (!conv && it < maxit) {
    const double rho1 = vdot(n, rh, r);
    if (rho1 == 0.0) break;
    const double beta = (rho1 / rho) * (alpha / omega);
    for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    amul(actx, p, tmp); pc(pctx, tmp, v);
    alpha = rho1 / vdot(n, rh, v);
    for (int64_t i = 0; i < n; ++i) sv[i] = r[i] - alpha * v[i];
    amul(actx, sv, tmp); pc(pctx, tmp, t);
    const double tt = vdot(n, t, t);
    omega = (tt == 0.0) ? 0.0 : vdot(n, t, sv) / tt;
    for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i] + omega * sv[i]; r[i] = sv[i] - omega * t[i]; }
    rho = rho1;
    ++it;
    res = vnorm(n, r);
    if (res <= rtol * bnorm) conv = 1;
    if (omega == 0.0) break;
  }
  *its = it; *rnorm = res;
  free(r); free(rh); free(p); free(v); free(sv); free(t); free(tmp);
  return conv ? 0 : 1;
}

typedef struct { int64_t n; int k; const double *a; } band_ctx;
typedef struct { int n; const int *ia, *ja; const double *a; } csr_ctx;
static void band_amul(void *c, const double *x, double *y) { band_ctx *b = (band_ctx *)c; orc_band_mult(b->n, b->k, b->a, x, y); }
static void band_pc(void *c, const double *x, double *y) {
  band_ctx *b = (band_ctx *)c;
  if (y != x) memcpy(y, x, sizeof(double) * (size_t)b->n);
  band_solve_range(0, b->n, b->k, b->a, y);
}
static void csr_amul(void *c, const double *x, double *y) {
  csr_ctx *m = (csr_ctx *)c;
  for (int i = 0; i < m->n; ++i) { double s = 0; for (int q = m->ia[i]; q < m->ia[i + 1]; ++q) s += m->a[q] * x[m->ja[q]]; y[i] = s; }
}
int orc_krylov_band(int64_t n, int k, const double *a, const double *lu, int method, int restart,
                    double rtol, int maxit, const double *b, double *x, int *its, double *rnorm) {
  band_ctx A = {n, k, a}, M = {n, k, lu};
  if (method == 0) return orc_gmres(n, band_amul, &A, band_pc, &M, b, x, restart, rtol, maxit, its, rnorm);
  return orc_bicgstab(n, band_amul, &A, band_pc, &M, b, x, rtol, maxit, its, rnorm);
}
int orc_krylov_csr_band(int n, const int *ia, const int *ja, const double *a, int k, const double *lu,
                        int method, int restart, double rtol, int maxit, const double *b, double *x,
                        int *its, double *rnorm) {
  csr_ctx A = {n, ia, ja, a}; band_ctx M = {n, k, lu};
  if (method == 0) return orc_gmres(n, csr_amul, &A, band_pc, &M, b, x, restart, rtol, maxit, its, rnorm);
  return orc_bicgstab(n, csr_amul, &A, band_pc, &M, b, x, rtol, maxit, its, rnorm);
}
