/*
 * spike_oracle.h -- CPU ORACLE for the spike-petsc hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (spike_petsc_b200/) never links or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - band selection / extraction, MatPermute / VecPermute semantics, AWBM: restated from the
 *     reference sources cited per function; pinned by the reference's own 3x3 input
 *     (src/wbm.c:483-497) and the probe answers recorded in SURVEY.md 8c (tests/golden/).
 *   - MC64 (WBM): not restated -- the reference file src/hslmc64.c is compiled as-is into
 *     oracle/_ref/ (see oracle/Makefile) and called through the wrapper convention of
 *     src/petsc_mat_wbm.c:20-58.
 *   - banded LU / solve, SPIKE, Krylov: the reference delegates these to PETSc (un-vendored,
 *     src/matbanded.c:178,190; src/kspreorder.c:124) and ships no expected outputs
 *     => "parity unpinned" for that arithmetic; pinned here against scipy/LAPACK instead.
 */
#ifndef SPIKE_ORACLE_H
#define SPIKE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs ------------------------------------------------------ */
uint64_t orc_splitmix64(uint64_t z);
double   orc_u01(uint64_t seed, uint64_t counter);
/* rows layout: a[i*(2k+1) + (j-i+k)], out-of-range entries are 0 */
void orc_gen_band(int64_t n, int k, uint64_t seed, double delta, double *a);
void orc_gen_vec(int64_t n, uint64_t seed, double *u);

/* ---- reference restatements --------------------------------------------------------------- */
/* src/matbanded.c:38-56,104-105: choose k, return normB/normA.  CSR is 0-based. */
int orc_band_select(int n, const int *ia, const int *ja, const double *a, int kmax, double frac,
                    int *k_out, double *frac_out);
/* src/matbanded.c:84-99: copy entries with |c-r|<=k keeping column order; returns nnz written. */
int64_t orc_band_extract_csr(int n, const int *ia, const int *ja, const double *a, int k,
                             int *ib, int *jb, double *b);
/* CSR -> rows layout band (entries with |c-r|<=k; duplicates are summed). */
void orc_csr_to_band(int n, const int *ia, const int *ja, const double *a, int k, double *band);
/* PETSc MatPermute semantics used at src/kspreorder.c:20: B(i,j) = A(rowp[i], colp[j]); rows sorted. */
int orc_mat_permute_csr(int n, const int *ia, const int *ja, const double *a, const int *rowp,
                        const int *colp, int *ib, int *jb, double *b);
/* PETSc VecPermute semantics used at src/kspreorder.c:122-127 (in place). */
void orc_vec_permute(int n, double *x, const int *idx, int inverse);
/* src/petsc_mat_awbm.c:65-205.  Returns 0 or a negative error; permR = p, match[] optional. */
int orc_awbm(int n, const int *ia, const int *ja, const double *a, int *permR, int *match_out);

/* ---- exact banded solve: the "reference CPU MATBANDED path" stand-in ---------------------- */
/* in-place no-pivot LU with diagonal boosting; returns number of boosted pivots */
int64_t orc_band_lu(int64_t n, int k, double *a, double boost);
void    orc_band_solve(int64_t n, int k, const double *lu, double *x, int nrhs, int64_t ldx);
void    orc_band_mult(int64_t n, int k, const double *a, const double *x, double *y);

/* ---- CPU SPIKE (truncated), partition-parallel with OpenMP -------------------------------- */
typedef struct orc_spike orc_spike;
/* nparts partitions with boundaries at multiples of `align` rows; tip_rows<=0 => exact UL over the
 * whole partition and full second sweep; >0 => windowed UL / truncated corrections of tip_rows. */
orc_spike *orc_spike_create(int64_t n, int k, int nparts, int align, int64_t tip_rows, double boost);
void       orc_spike_destroy(orc_spike *s);
/* a is the rows-layout band; it is factored IN PLACE (coupling blocks keep their A values). */
int64_t    orc_spike_factor(orc_spike *s, double *a, int nthreads);
void       orc_spike_solve(orc_spike *s, const double *a_lu, const double *b, double *x, int nthreads);
/* introspection for tip parity: V^(b) of interface i and W^(t) of interface i (kxk row-major) */
const double *orc_spike_vb(const orc_spike *s, int iface);
const double *orc_spike_wt(const orc_spike *s, int iface);
int64_t       orc_spike_part_start(const orc_spike *s, int part);

/* ---- Krylov (left-preconditioned, PETSc defaults) ------------------------------------------ */
typedef void (*orc_apply_fn)(void *ctx, const double *x, double *y);
/* GMRES(restart), classical Gram-Schmidt, preconditioned residual norm, rtol relative to ||M^-1 b|| */
int orc_gmres(int64_t n, orc_apply_fn amul, void *actx, orc_apply_fn pc, void *pctx, const double *b,
              double *x, int restart, double rtol, int maxit, int *its, double *rnorm);
int orc_bicgstab(int64_t n, orc_apply_fn amul, void *actx, orc_apply_fn pc, void *pctx,
                 const double *b, double *x, double rtol, int maxit, int *its, double *rnorm);
/* convenience drivers used from Python: A = band (rows layout), M^-1 = exact band LU solve */
int orc_krylov_band(int64_t n, int k, const double *a, const double *lu, int method, int restart,
                    double rtol, int maxit, const double *b, double *x, int *its, double *rnorm);
/* A = CSR, M^-1 = exact band LU solve of half-bandwidth k */
int orc_krylov_csr_band(int n, const int *ia, const int *ja, const double *a, int k, const double *lu,
                        int method, int restart, double rtol, int maxit, const double *b, double *x,
                        int *its, double *rnorm);

int orc_num_threads(void);
#ifdef __cplusplus
}
#endif
#endif
