/*
 * kspreorder.c -- KSPREORDER ("reorder"): solve of a non-symmetrically permuted system.
 *
 * Mirrors /root/reference/src/kspreorder.c: setup computes (rorder, corder) = MatGetOrdering(M, type)
 * (:19) with -mat_ordering_type under the KSP's prefix (:146); solve permutes x by corder and b by
 * rorder IN PLACE, runs the inner Krylov solve, copies the converged reason and un-permutes
 * (:122-127).  The orderings stay host code registered by name (wbm / awbm / fiedler,
 * src/testbed2.c:66-68); what moves to the GPU is everything they feed:
 *   MatPermute (:20-21)          -> fused into spk_set_band_csr / spk_set_operator_csr (gather)
 *   VecPermute x4 (:122-127)     -> spk_permute (gather / scatter kernels)
 *   inner KSPSolve (:124)        -> spk_krylov (left-preconditioned GMRES / BiCGStab on the device,
 *                                   M^-1 = SPIKE solve of the PCBANDED band)
 * Inner options use the "reorder_" prefix (:218-221): -reorder_ksp_type gmres|bcgs, -reorder_ksp_rtol,
 * -reorder_ksp_max_it, -reorder_ksp_gmres_restart, -reorder_pc_type banded, -reorder_pc_banded_kmax/frac.
 */
#include "petsc_access.h"
#include "../../include/spike_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

PetscErrorCode PCBandedSetPermutation_Private(PC pc, const PetscInt *rowperm, const PetscInt *colperm);
spk_ctx *PCBandedGetContext_Private(PC pc);

typedef struct {
  PC   pc;               /* the embedded (inner) preconditioner: PCBANDED */
  char ordertype[256];   /* The type of ordering */
  IS   rorder, corder;   /* The row and column orderings */
  int  method, restart;  /* inner KSP type */
} KSP_Reorder;

static PetscErrorCode KSPSetFromOptions_Reorder(KSP ksp) {   /* :134-151 */
  KSP_Reorder *r = (KSP_Reorder *)ksp->data;
  char inner[144], tname[64] = "";
  PetscBool flg;
  snprintf(r->ordertype, sizeof r->ordertype, "natural");   /* MATORDERINGNATURAL, :144 */
  PetscOptionsGetString(SPK_PREFIX(ksp), "-mat_ordering_type", r->ordertype, sizeof r->ordertype, NULL);
  snprintf(inner, sizeof inner, "%sreorder_", SPK_PREFIX(ksp));
  PetscOptionsGetString(inner, "-ksp_type", tname, sizeof tname, &flg);
  if (flg) {
    if (!strcmp(tname, "gmres")) r->method = SPK_KSP_GMRES;
    else if (!strcmp(tname, "bcgs")) r->method = SPK_KSP_BCGS;
    else SPK_ERR(PETSC_ERR_SUP, "inner KSP type %s not available on the device (gmres, bcgs)", tname);
  }
  PetscOptionsGetReal(inner, "-ksp_rtol", &ksp->rtol, NULL);
  PetscOptionsGetInt(inner, "-ksp_max_it", &ksp->max_it, NULL);
  PetscOptionsGetInt(inner, "-ksp_gmres_restart", &r->restart, NULL);
  PetscOptionsGetString(inner, "-pc_type", tname, sizeof tname, &flg);
  if (flg && strcmp(tname, "banded")) SPK_ERR(PETSC_ERR_SUP, "inner PC type %s not available (banded)", tname);
  PCSetOptionsPrefix(r->pc, inner);
  return PCSetFromOptions(r->pc);
}

static PetscErrorCode KSPSetUp_Reorder(KSP ksp) {            /* :11-28 */
  KSP_Reorder *r = (KSP_Reorder *)ksp->data;
  PetscErrorCode ierr;
  Mat A = ksp->A, M = ksp->M ? ksp->M : ksp->A;
  PetscInt n, nr, nc; const PetscInt *ai, *aj, *ridx, *cidx; const PetscScalar *aa;
  if (!A) SPK_ERR(PETSC_ERR_ARG_WRONGSTATE, "KSPREORDER: no operators");
  if (r->rorder) { ISDestroy(&r->rorder); ISDestroy(&r->corder); }   /* the reference leaks these on re-setup */
  ierr = MatGetOrdering(M, r->ordertype, &r->rorder, &r->corder);CHKERRQ(ierr);     /* :19 */
  ierr = SpkMatGetCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  ierr = SpkISGetIndices(r->rorder, &nr, &ridx);CHKERRQ(ierr);
  ierr = SpkISGetIndices(r->corder, &nc, &cidx);CHKERRQ(ierr);
  if (nr != n || nc != n) SPK_ERR(PETSC_ERR_ARG_OUTOFRANGE, "ordering has wrong length");
  /* PM = MatPermute(M, rorder, corder) (:20) is fused into the band extraction of the inner PC ...  */
  ierr = PCSetOperators(r->pc, A, M);CHKERRQ(ierr);
  ierr = PCBandedSetPermutation_Private(r->pc, ridx, cidx);CHKERRQ(ierr);
  ierr = PCSetUp(r->pc);CHKERRQ(ierr);                                              /* KSPSetUp(inner), :24 */
  /* ... and PA = MatPermute(A, rorder, corder) (:21) into the device copy of the Krylov operator */
  spk_ctx *ctx = PCBandedGetContext_Private(r->pc);
  if (spk_set_operator_csr(ctx, n, ai, aj, aa, ridx, cidx))
    SPK_ERR(PETSC_ERR_LIB, "KSPREORDER: %s", spk_last_error(ctx));
  ierr = SpkMatRestoreCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  return 0;   /* (the index arrays stay borrowed by the inner PC until the ISs are destroyed) */
}

static PetscErrorCode KSPSolve_Reorder(KSP ksp) {            /* :113-128 */
  KSP_Reorder *r = (KSP_Reorder *)ksp->data;
  Vec x = ksp->vec_sol, b = ksp->vec_rhs;
  spk_ctx *ctx = PCBandedGetContext_Private(r->pc);
  int its = 0, conv = 0; double rn = 0.0;
  PetscInt n, nr; PetscScalar *xa, *ba; const PetscInt *ridx, *cidx; PetscErrorCode ierr;
  if (!ctx) SPK_ERR(PETSC_ERR_ARG_WRONGSTATE, "KSPREORDER: solve before setup");
  ierr = SpkVecGetArray(x, &n, &xa);CHKERRQ(ierr);
  ierr = SpkVecGetArray(b, &n, &ba);CHKERRQ(ierr);
  ierr = SpkISGetIndices(r->rorder, &nr, &ridx);CHKERRQ(ierr);
  ierr = SpkISGetIndices(r->corder, &nr, &cidx);CHKERRQ(ierr);
  if (spk_permute(ctx, cidx, 0, xa, n)) SPK_ERR(PETSC_ERR_LIB, "%s", spk_last_error(ctx));   /* :122 */
  if (spk_permute(ctx, ridx, 0, ba, n)) SPK_ERR(PETSC_ERR_LIB, "%s", spk_last_error(ctx));   /* :123 */
  if (spk_krylov(ctx, r->method, r->restart, ksp->rtol, ksp->max_it, ba, xa, &its, &rn, &conv))           /* :124 */
    SPK_ERR(PETSC_ERR_LIB, "KSPREORDER: %s", spk_last_error(ctx));
  ksp->its = its; ksp->rnorm = rn;
  ksp->reason = conv ? 2 /* KSP_CONVERGED_RTOL */ : -3 /* KSP_DIVERGED_ITS */;                                /* :125 */
  if (spk_permute(ctx, cidx, 1, xa, n)) SPK_ERR(PETSC_ERR_LIB, "%s", spk_last_error(ctx));   /* :126 */
  if (spk_permute(ctx, ridx, 1, ba, n)) SPK_ERR(PETSC_ERR_LIB, "%s", spk_last_error(ctx));   /* :127 */
  ierr = SpkVecRestoreArray(x, &xa);CHKERRQ(ierr);
  ierr = SpkVecRestoreArray(b, &ba);CHKERRQ(ierr);
  return 0;
}

static PetscErrorCode KSPView_Reorder(KSP ksp, char *buf, size_t len) {   /* :155-170 */
  KSP_Reorder *r = (KSP_Reorder *)ksp->data;
  int n = snprintf(buf, len, "  reordering type = %s\n", r->ordertype);
  if (n > 0 && (size_t)n < len && r->pc->ops->view) r->pc->ops->view(r->pc, buf + n, len - (size_t)n);
  return 0;
}
static PetscErrorCode KSPDestroy_Reorder(KSP ksp) {          /* :174-185 */
  KSP_Reorder *r = (KSP_Reorder *)ksp->data;
  ISDestroy(&r->rorder); ISDestroy(&r->corder);
  PCDestroy(&r->pc);
  free(r); ksp->data = NULL;
  return 0;
}

/* the embedded preconditioner (the reference reaches it through the "reorder_" options prefix only) */
PetscErrorCode KSPReorderGetPC(KSP ksp, PC *pc) { *pc = ((KSP_Reorder *)ksp->data)->pc; return 0; }

PetscErrorCode KSPCreate_Reorder(KSP ksp) {                  /* :197-223 */
  KSP_Reorder *r = (KSP_Reorder *)calloc(1, sizeof(*r));
  PetscErrorCode ierr;
  ksp->data = (void *)r;
  r->method = SPK_KSP_GMRES; r->restart = 30;
  snprintf(r->ordertype, sizeof r->ordertype, "natural");
  ksp->ops->setup          = KSPSetUp_Reorder;
  ksp->ops->solve          = KSPSolve_Reorder;
  ksp->ops->destroy        = KSPDestroy_Reorder;
  ksp->ops->view           = KSPView_Reorder;
  ksp->ops->setfromoptions = KSPSetFromOptions_Reorder;
  ierr = PCCreate(&r->pc);CHKERRQ(ierr);
  ierr = PCCreate_Banded(r->pc);CHKERRQ(ierr);
  { char inner[144]; snprintf(inner, sizeof inner, "%sreorder_", SPK_PREFIX(ksp)); PCSetOptionsPrefix(r->pc, inner); }   /* :218-221 */
  return 0;
}
