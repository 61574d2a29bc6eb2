/*
 * petsc_access.h -- the ONLY place where the host glue touches the inside of a Mat / Vec / IS or an object's
 * options prefix.  Two implementations of the same accessors:
 *   - default: the shim of petscshim.h (this image has no PETSc), plain struct fields;
 *   - -DHAVE_PETSC: the real PETSc calls (MatGetRowIJ + MatSeqAIJGetArrayRead, VecGetArray, ISGetIndices,
 *     PetscObject prefix), PETSc >= 3.5 spelling.  That branch cannot be compiled in this image and is therefore
 *     compile-UNTESTED; it is the binding INTEGRATION.md describes, kept next to the tested one so the glue
 *     sources (pcbanded.c, kspreorder.c, matbanded_type.c) contain no shim field access of Mat/Vec/IS.
 * PC / KSP private fields used by the glue (->data, ->ops, ->pmat, ->setupcalled, ->vec_rhs, ->vec_sol, ->its,
 * ->reason, ->rnorm, ->rtol, ->max_it) carry the same names in petsc/private/pcimpl.h and kspimpl.h.
 */
#ifndef SPK_PETSC_ACCESS_H
#define SPK_PETSC_ACCESS_H
#include "spike_petsc.h"

#ifdef HAVE_PETSC
#include <petsc/private/pcimpl.h>
#include <petsc/private/kspimpl.h>
#define SPK_ERR(code, ...) SETERRQ(PETSC_COMM_SELF, code, __VA_ARGS__)
#define SPK_PREFIX(obj) (((PetscObject)(obj))->prefix ? ((PetscObject)(obj))->prefix : "")
static inline PetscErrorCode SpkMatGetCSR(Mat A, PetscInt *n, const PetscInt **ia, const PetscInt **ja, const PetscScalar **a) {
  PetscBool done; PetscErrorCode ierr;
  ierr = MatGetRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, n, ia, ja, &done);CHKERRQ(ierr);
  if (!done) SPK_ERR(PETSC_ERR_SUP, "MatGetRowIJ not available for this Mat type (SeqAIJ expected)");
  ierr = MatSeqAIJGetArrayRead(A, a);CHKERRQ(ierr);
  return 0;
}
static inline PetscErrorCode SpkMatRestoreCSR(Mat A, PetscInt *n, const PetscInt **ia, const PetscInt **ja, const PetscScalar **a) {
  PetscBool done; PetscErrorCode ierr;
  ierr = MatSeqAIJRestoreArrayRead(A, a);CHKERRQ(ierr);
  ierr = MatRestoreRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, n, ia, ja, &done);CHKERRQ(ierr);
  return 0;
}
static inline PetscErrorCode SpkVecGetArray(Vec v, PetscInt *n, PetscScalar **a) { PetscErrorCode ierr = VecGetLocalSize(v, n);CHKERRQ(ierr); return VecGetArray(v, a); }
static inline PetscErrorCode SpkVecRestoreArray(Vec v, PetscScalar **a) { return VecRestoreArray(v, a); }
static inline PetscErrorCode SpkISGetIndices(IS is, PetscInt *n, const PetscInt **idx) { PetscErrorCode ierr = ISGetLocalSize(is, n);CHKERRQ(ierr); return ISGetIndices(is, idx); }
static inline PetscErrorCode SpkISRestoreIndices(IS is, const PetscInt **idx) { return ISRestoreIndices(is, idx); }
#else
#define SPK_ERR(code, ...) SETERRQ(code, __VA_ARGS__)
#define SPK_PREFIX(obj) ((obj)->prefix)
static inline PetscErrorCode SpkMatGetCSR(Mat A, PetscInt *n, const PetscInt **ia, const PetscInt **ja, const PetscScalar **a) {
  if (!A->i || !A->j) SPK_ERR(PETSC_ERR_SUP, "Mat type %s has no CSR arrays (SeqAIJ expected)", A->type);
  *n = A->n; *ia = A->i; *ja = A->j; *a = A->a; return 0;
}
static inline PetscErrorCode SpkMatRestoreCSR(Mat A, PetscInt *n, const PetscInt **ia, const PetscInt **ja, const PetscScalar **a) {
  (void)A; (void)n; *ia = NULL; *ja = NULL; *a = NULL; return 0;
}
static inline PetscErrorCode SpkVecGetArray(Vec v, PetscInt *n, PetscScalar **a) { *n = v->n; *a = v->a; return 0; }
static inline PetscErrorCode SpkVecRestoreArray(Vec v, PetscScalar **a) { (void)v; *a = NULL; return 0; }
static inline PetscErrorCode SpkISGetIndices(IS is, PetscInt *n, const PetscInt **idx) { *n = is->n; *idx = is->idx; return 0; }
static inline PetscErrorCode SpkISRestoreIndices(IS is, const PetscInt **idx) { (void)is; *idx = NULL; return 0; }
#endif
#endif
