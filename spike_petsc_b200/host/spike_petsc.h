/*
 * spike_petsc.h -- host-side mirror of the reference's plugin surface, implemented on the C ABI of
 * include/spike_b200.h.  Same names, argument meaning and error behaviour as the reference:
 *   MatCreateSubMatrixBanded   /root/reference/src/matbanded.h:5, src/matbanded.c:22-107
 *   PCCreate_Banded ("banded") /root/reference/src/matbanded.c:251-283 (+ ops :120-233, setters :305-343)
 *   KSPCreate_Reorder ("reorder") /root/reference/src/kspreorder.c:197-223 (+ ops :11-28,113-185)
 * Registration is by name exactly as src/testbed2.c:66-71 does it.
 */
#ifndef SPK_SPIKE_PETSC_H
#define SPK_SPIKE_PETSC_H
#ifdef HAVE_PETSC
#include <petscksp.h>
#else
#include "petscshim.h"
#endif
#ifdef __cplusplus
extern "C" {
#endif
PetscErrorCode MatCreateSubMatrixBanded(Mat A, PetscInt *kmax, PetscReal *frac, Mat *B);
PetscErrorCode PCCreate_Banded(PC pc);
PetscErrorCode PCBandedSetMaxHalfBandwith(PC pc, PetscInt kmax);   /* (sic) reference spelling, src/matbanded.c:305 */
PetscErrorCode PCBandedSetNormFraction(PC pc, PetscReal frac);
PetscErrorCode PCBandedGetInfo(PC pc, PetscInt *k, PetscReal *f, PetscInt *partitions, long long *boosted);
PetscErrorCode PCBandedGetSetupCount(PC pc, PetscInt *nsetup);
/* error of a probe solve against the exact band solve measured at setup (-spike_verify 1, the default), and whether
 * the PC fell back to the exact single-partition mode because it exceeded -spike_verify_tol (default 1e-6) */
PetscErrorCode PCBandedGetApplyError(PC pc, PetscReal *err, PetscInt *exact_fallback);
PetscErrorCode KSPCreate_Reorder(KSP ksp);
PetscErrorCode KSPReorderGetPC(KSP ksp, PC *pc);
/* MATBANDED Mat type (matbanded_type.c): MatRegister("banded", MatCreate_Banded); ops mult / lufactor / solve / matsolve /
 * getdiagonal / view / destroy on the GPU band.  MatCreateBanded = MatCreate + MatSetType + MatBandedSetFromAIJ. */
PetscErrorCode MatCreate_Banded(Mat B);
PetscErrorCode MatBandedSetFromAIJ(Mat B, Mat A, PetscInt *kmax, PetscReal *frac);
PetscErrorCode MatCreateBanded(Mat A, PetscInt *kmax, PetscReal *frac, Mat *B);
PetscErrorCode MatBandedGetInfo(Mat B, PetscInt *k, PetscReal *f, PetscInt *nfactor);
/* host orderings (ordering.c): reverse Cuthill-McKee and the deterministic stand-in for the reference's MC73-based
 * "fiedler" ordering (src/petsc_mat_fiedler.c:11-58; MC73 is proprietary and absent).  perm[new] = old. */
int SpkOrderingRCM(PetscInt n, const PetscInt *ai, const PetscInt *aj, PetscInt *perm);
PetscErrorCode MatGetOrdering_RCM(Mat A, const char *type, IS *row, IS *col);
PetscErrorCode MatGetOrdering_Fiedler(Mat A, const char *type, IS *row, IS *col);
/* "awbm" with the matching computed on the GPU (spk_awbm_csr), bit-identical to src/petsc_mat_awbm.c:42-225 */
PetscErrorCode MatGetOrdering_AWBM(Mat A, const char *type, IS *row, IS *col);
/* on-disk formats of the reference's drivers (matio.c): PETSc binary Mat/Vec (MatLoad, src/testbed2.c:93-96),
 * MatrixMarket export (src/wbm.c:520-523).  0-based CSR, arrays malloc'd for the caller (SpkFree). */
int SpkMatLoadBinary(const char *path, int *rows, int *cols, int **ia, int **ja, double **a);
int SpkMatWriteBinary(const char *path, int rows, int cols, const int *ia, const int *ja, const double *a);
int SpkVecLoadBinary(const char *path, int *n, double **v);
int SpkVecWriteBinary(const char *path, int n, const double *v);
int SpkMatWriteMatrixMarket(const char *path, int rows, int cols, const int *ia, const int *ja, const double *a, int digits);
void SpkFree(void *p);
#ifdef __cplusplus
}
#endif
#endif
