/*
 * pcbanded.c -- PCBANDED ("banded") and MatCreateSubMatrixBanded on the B200 SPIKE engine.
 *
 * Mirrors /root/reference/src/matbanded.c: same type name, ops-table slots, option names
 * (-pc_banded_kmax, -pc_banded_frac), defaults (kmax 50, frac 0.95; :261-262), view text (:205) and
 * setup/apply split.  Where the reference builds an AIJ band B and hands it to an inner PETSc PC
 * (:174-178) this PC hands the operator to spk_set_band_csr (band selection + extraction, fused with
 * the KSPREORDER permutation when one is attached) and spk_factor; PCApply (:190) is spk_solve.
 * Nothing here computes on the CPU except MatCreateSubMatrixBanded, which the reference exports as
 * a stand-alone host utility (src/matbanded.h:5) and which is restated for API completeness.
 */
#include "petsc_access.h"
#include "../../include/spike_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- MatCreateSubMatrixBanded, /root/reference/src/matbanded.c:22-107 (host utility) ---------- */
PetscErrorCode MatCreateSubMatrixBanded(Mat A, PetscInt *kmax, PetscReal *frac, Mat *B) {
  PetscInt n; const PetscInt *ai, *aj; const PetscScalar *aa;
  PetscErrorCode ierr = SpkMatGetCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  PetscReal *w = (PetscReal *)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscReal));
  PetscReal normA = 0.0, normB = 0.0;
  PetscInt r, c, k;
  for (r = 0; r < n; ++r)                       /* :38-49 */
    for (c = ai[r]; c < ai[r + 1]; ++c) { w[abs(r - aj[c])] += fabs(aa[c]); normA += fabs(aa[c]); }
  for (k = 0; k < *kmax; ++k) {                 /* :53-56; the reference has no k<n guard (reads past w) */
    if (k >= n) { free(w); SPK_ERR(PETSC_ERR_ARG_OUTOFRANGE, "kmax %d exceeds matrix order %d before the norm fraction is reached", *kmax, n); }
    normB += w[k];
    if (normB >= (*frac) * normA) break;
  }
  free(w);
  PetscInt nnz = 0;                             /* :65-79 count, :84-99 copy (column order preserved) */
  for (r = 0; r < n; ++r) for (c = ai[r]; c < ai[r + 1]; ++c) if (abs(aj[c] - r) <= k) ++nnz;
  PetscInt *bi = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n + 1));
  PetscInt *bj = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(nnz > 0 ? nnz : 1));
  PetscScalar *ba = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(nnz > 0 ? nnz : 1));
  nnz = 0; bi[0] = 0;
  for (r = 0; r < n; ++r) {
    for (c = ai[r]; c < ai[r + 1]; ++c) { if (abs(aj[c] - r) > k) continue; bj[nnz] = aj[c]; ba[nnz] = aa[c]; ++nnz; }
    bi[r + 1] = nnz;
  }
  ierr = MatCreateSeqAIJWithArrays(n, bi, bj, ba, B);
  free(bi); free(bj); free(ba);
  CHKERRQ(ierr);
  ierr = SpkMatRestoreCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  *kmax = k;                                    /* :104-105 */
  *frac = normB / normA;
  return 0;
}

/* ---- PC_Banded, /root/reference/src/matbanded.c:111-116 ---------------------------------------- */
typedef struct {
  PetscInt  kmax, k;     /* maximum and actual half-bandwidth */
  PetscReal frac, f;     /* norm fraction limit and actual */
  spk_ctx  *ctx;         /* replaces {Mat B; PC pc;}: the band lives on the GPU, factored in place */
  PetscInt  partitions, tip_tiles;
  const PetscInt *rowperm, *colperm;   /* borrowed from KSPREORDER: B = band(pmat(rowperm, colperm)) */
  PetscInt  nsetup;                    /* factorisations performed (diagnostics / tests) */
  PetscInt  verify;                    /* -spike_verify: probe solve after the factorisation (default 1) */
  PetscReal verify_tol;                /* -spike_verify_tol: above it the PC falls back to the exact mode (1 partition) */
  PetscReal apply_err;                 /* ||x - v|| / ||v|| of the probe solve, -1 = not measured */
  PetscInt  exact_fallback;            /* the fallback was taken */
} PC_Banded;

static PetscErrorCode PCReset_Banded(PC pc) {          /* :120-129 */
  PC_Banded *b = (PC_Banded *)pc->data;
  if (b->ctx) spk_destroy(&b->ctx);
  return 0;
}
static PetscErrorCode PCDestroy_Banded(PC pc) {        /* :133-145 */
  PetscErrorCode ierr = PCReset_Banded(pc); CHKERRQ(ierr);
  free(pc->data); pc->data = NULL;
  return 0;
}
static PetscErrorCode PCSetFromOptions_Banded(PC pc) { /* :149-161 */
  PC_Banded *b = (PC_Banded *)pc->data;
  PetscOptionsGetInt(SPK_PREFIX(pc), "-pc_banded_kmax", &b->kmax, NULL);
  PetscOptionsGetReal(SPK_PREFIX(pc), "-pc_banded_frac", &b->frac, NULL);
  /* engine knobs (new): inner-object prefix "banded_" like the reference's embedded PC (:278-281) */
  char inner[192]; snprintf(inner, sizeof inner, "%sbanded_", SPK_PREFIX(pc));
  PetscOptionsGetInt(inner, "-spike_partitions", &b->partitions, NULL);
  PetscOptionsGetInt(inner, "-spike_tip_tiles", &b->tip_tiles, NULL);
  PetscOptionsGetInt(inner, "-spike_verify", &b->verify, NULL);
  PetscOptionsGetReal(inner, "-spike_verify_tol", &b->verify_tol, NULL);
  return 0;
}
/* First call (setupcalled == 0, :171): choose k, extract the band, factor.  A later call means PETSc saw the operator
 * change (PCSetUp returns early otherwise); the reference then re-runs PCSetUp(inner) on its stale B (:178) -- here
 * the band is extracted from the current pmat again and refactored, which is what a changed operator needs. */
static PetscErrorCode PCSetUp_Banded(PC pc) {          /* :165-180 */
  PC_Banded *b = (PC_Banded *)pc->data;
  PetscInt n; const PetscInt *ai, *aj; const PetscScalar *aa;
  PetscErrorCode ierr;
  if (!pc->pmat) SPK_ERR(PETSC_ERR_ARG_WRONGSTATE, "PCBANDED: no preconditioner matrix set");
  /* The reference's inner PC is an exact LU of B (:178).  The truncated SPIKE is exact only for bands whose spikes decay
   * inside the truncation window (diagonally dominant ones), and a reordered band need not be one: after factoring,
   * one probe solve through the kept unfactored band measures ||B^-1 b - x||; above -spike_verify_tol the PC refactors
   * in the exact mode (-spike_partitions 1: one partition, nothing truncated) and says so in PCView. */
  b->apply_err = -1.0; b->exact_fallback = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (b->ctx) spk_destroy(&b->ctx);
    spk_opts o; spk_default_opts(&o);
    o.partitions = attempt ? 1 : b->partitions; o.tip_tiles = attempt ? -1 : b->tip_tiles; o.mem = SPK_MEM_HOST;
    if (spk_create(&b->ctx, &o)) SPK_ERR(PETSC_ERR_LIB, "PCBANDED: %s", spk_last_error(NULL));
    if (b->verify) spk_keep_original(b->ctx, 1);
    b->k = b->kmax; b->f = b->frac;                      /* :172-173 */
    int k = b->k; double f = b->f;
    /* MatCreateSubMatrixBanded(pc->pmat, &b->k, &b->f, &b->B) (:174), on the (permuted) operator */
    ierr = SpkMatGetCSR(pc->pmat, &n, &ai, &aj, &aa);CHKERRQ(ierr);
    if (spk_set_band_csr(b->ctx, n, ai, aj, aa, b->rowperm, b->colperm, &k, &f))
      SPK_ERR(PETSC_ERR_LIB, "PCBANDED: %s", spk_last_error(b->ctx));
    ierr = SpkMatRestoreCSR(pc->pmat, &n, &ai, &aj, &aa);CHKERRQ(ierr);
    b->k = k; b->f = f;
    /* PCSetUp(b->pc) (:178): the SPIKE factorisation */
    if (spk_factor(b->ctx)) SPK_ERR(PETSC_ERR_LIB, "PCBANDED: %s", spk_last_error(b->ctx));
    if (!b->verify) break;
    double e = 0.0;
    if (spk_check(b->ctx, &e)) SPK_ERR(PETSC_ERR_LIB, "PCBANDED: %s", spk_last_error(b->ctx));
    b->apply_err = e;
    if (attempt == 1) { b->exact_fallback = 1; break; }
    if (e <= b->verify_tol) break;
    spk_info info; memset(&info, 0, sizeof info); spk_view(b->ctx, &info);
    if (info.partitions <= 1) break;                     /* already exact: the error is the LU's own */
  }
  b->nsetup++;
  return 0;
}
static PetscErrorCode PCApply_Banded(PC pc, Vec x, Vec y) {  /* :184-192 */
  PC_Banded *b = (PC_Banded *)pc->data;
  PetscInt n; PetscScalar *xa, *ya; PetscErrorCode ierr;
  if (!b->ctx) SPK_ERR(PETSC_ERR_ARG_WRONGSTATE, "PCBANDED: apply before setup");
  ierr = SpkVecGetArray(x, &n, &xa);CHKERRQ(ierr);
  ierr = SpkVecGetArray(y, &n, &ya);CHKERRQ(ierr);
  if (spk_solve(b->ctx, xa, ya, 1)) SPK_ERR(PETSC_ERR_LIB, "PCBANDED: %s", spk_last_error(b->ctx));
  ierr = SpkVecRestoreArray(x, &xa);CHKERRQ(ierr);
  ierr = SpkVecRestoreArray(y, &ya);CHKERRQ(ierr);
  return 0;
}
static PetscErrorCode PCView_Banded(PC pc, char *buf, size_t len) {  /* :196-211 */
  PC_Banded *b = (PC_Banded *)pc->data;
  spk_info info; memset(&info, 0, sizeof info);
  if (b->ctx) spk_view(b->ctx, &info);
  int w = snprintf(buf, len, "  Banded: k = %d (%d max), frac = %g (%g max)\n    SPIKE (B200): partitions = %d, tip window = %d tiles, boosted pivots = %lld\n",
           b->k, b->kmax, b->f, b->frac, info.partitions, info.tip_tiles, (long long)info.boosted_pivots);
  if (w > 0 && (size_t)w < len && b->apply_err >= 0.0)
    snprintf(buf + w, len - (size_t)w, "    probe solve vs the exact band solve: relative error %.3e%s\n", b->apply_err,
             b->exact_fallback ? " (truncated SPIKE rejected: refactored with 1 partition, the exact mode)" : "");
  return 0;
}

PetscErrorCode PCCreate_Banded(PC pc) {                /* :251-283 */
  PC_Banded *b = (PC_Banded *)calloc(1, sizeof(*b));
  pc->data = (void *)b;
  b->kmax = 50;
  b->frac = 0.95;
  b->verify = 1; b->verify_tol = 1e-6; b->apply_err = -1.0;
  pc->ops->apply          = PCApply_Banded;
  pc->ops->applytranspose = NULL;
  pc->ops->setup          = PCSetUp_Banded;
  pc->ops->reset          = PCReset_Banded;
  pc->ops->destroy        = PCDestroy_Banded;
  pc->ops->setfromoptions = PCSetFromOptions_Banded;
  pc->ops->view           = PCView_Banded;
  return 0;
}
/* setters, :305-343.  (In the reference both look up a misspelled composed name and silently do
 * nothing; here they take effect, which is the documented intent.) */
PetscErrorCode PCBandedSetMaxHalfBandwith(PC pc, PetscInt kmax) { ((PC_Banded *)pc->data)->kmax = kmax; return 0; }
PetscErrorCode PCBandedSetNormFraction(PC pc, PetscReal frac) { ((PC_Banded *)pc->data)->frac = frac; return 0; }
PetscErrorCode PCBandedGetInfo(PC pc, PetscInt *k, PetscReal *f, PetscInt *partitions, long long *boosted) {
  PC_Banded *b = (PC_Banded *)pc->data;
  spk_info info; memset(&info, 0, sizeof info);
  if (b->ctx) spk_view(b->ctx, &info);
  if (k) *k = b->k;
  if (f) *f = b->f;
  if (partitions) *partitions = info.partitions;
  if (boosted) *boosted = info.boosted_pivots;
  return 0;
}
/* used by KSPREORDER to fuse its MatPermute into the band extraction */
PetscErrorCode PCBandedSetPermutation_Private(PC pc, const PetscInt *rowperm, const PetscInt *colperm) {
  PC_Banded *b = (PC_Banded *)pc->data; b->rowperm = rowperm; b->colperm = colperm; return 0; }
spk_ctx *PCBandedGetContext_Private(PC pc) { return ((PC_Banded *)pc->data)->ctx; }
PetscErrorCode PCBandedGetApplyError(PC pc, PetscReal *err, PetscInt *exact_fallback) {
  PC_Banded *b = (PC_Banded *)pc->data; if (err) *err = b->apply_err; if (exact_fallback) *exact_fallback = b->exact_fallback; return 0; }
PetscErrorCode PCBandedGetSetupCount(PC pc, PetscInt *nsetup) { *nsetup = ((PC_Banded *)pc->data)->nsetup; return 0; }
