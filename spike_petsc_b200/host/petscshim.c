/* petscshim.c -- minimal object/option plumbing behind petscshim.h (not needed with real PETSc). */
#include "petscshim.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_msg[512];
const char *PetscLastErrorMessage(void) { return g_msg; }
void PetscSetErrorMessage(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_msg, sizeof g_msg, fmt, ap); va_end(ap); }

#define MAXOPT 64
static struct { char name[96], value[96]; } g_opt[MAXOPT];
static int g_nopt;
PetscErrorCode PetscOptionsClear(void) { g_nopt = 0; return 0; }
PetscErrorCode PetscOptionsSetValue(const char *name, const char *value) {
  for (int i = 0; i < g_nopt; ++i) if (!strcmp(g_opt[i].name, name)) { snprintf(g_opt[i].value, 96, "%s", value); return 0; }
  if (g_nopt >= MAXOPT) SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "too many options");
  snprintf(g_opt[g_nopt].name, 96, "%s", name); snprintf(g_opt[g_nopt].value, 96, "%s", value); ++g_nopt;
  return 0;
}
static const char *find_opt(const char *prefix, const char *name) {
  char full[192];
  snprintf(full, sizeof full, "-%s%s", prefix ? prefix : "", name + 1);  /* name starts with '-' */
  for (int i = 0; i < g_nopt; ++i) if (!strcmp(g_opt[i].name, full)) return g_opt[i].value;
  return NULL;
}
PetscErrorCode PetscOptionsGetInt(const char *p, const char *n, PetscInt *v, PetscBool *set) {
  const char *s = find_opt(p, n); if (set) *set = s ? PETSC_TRUE : PETSC_FALSE; if (s) *v = atoi(s); return 0; }
PetscErrorCode PetscOptionsGetReal(const char *p, const char *n, PetscReal *v, PetscBool *set) {
  const char *s = find_opt(p, n); if (set) *set = s ? PETSC_TRUE : PETSC_FALSE; if (s) *v = atof(s); return 0; }
PetscErrorCode PetscOptionsGetString(const char *p, const char *n, char *v, size_t len, PetscBool *set) {
  const char *s = find_opt(p, n); if (set) *set = s ? PETSC_TRUE : PETSC_FALSE; if (s) snprintf(v, len, "%s", s); return 0; }

static struct { char name[32]; MatOrderingFn fn; } g_ord[16];
static int g_nord;
static PetscErrorCode ordering_natural(Mat A, const char *t, IS *r, IS *c) {
  (void)t;
  PetscInt *idx = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)A->n);
  for (PetscInt i = 0; i < A->n; ++i) idx[i] = i;
  ISCreateGeneral(A->n, idx, r); ISCreateGeneral(A->n, idx, c); free(idx);
  return 0;
}
PetscErrorCode MatOrderingRegister(const char *name, MatOrderingFn fn) {
  for (int i = 0; i < g_nord; ++i) if (!strcmp(g_ord[i].name, name)) { g_ord[i].fn = fn; return 0; }
  if (g_nord >= 16) SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "too many orderings");
  snprintf(g_ord[g_nord].name, 32, "%s", name); g_ord[g_nord].fn = fn; ++g_nord; return 0;
}
PetscErrorCode MatGetOrdering(Mat A, const char *type, IS *row, IS *col) {
  if (!strcmp(type, "natural")) return ordering_natural(A, type, row, col);
  for (int i = 0; i < g_nord; ++i) if (!strcmp(g_ord[i].name, type)) return g_ord[i].fn(A, type, row, col);
  SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "Unknown ordering type %s", type);
}

PetscErrorCode MatCreateSeqAIJWithArrays(PetscInt n, const PetscInt *i, const PetscInt *j, const PetscScalar *a, Mat *A) {
  Mat m = (Mat)calloc(1, sizeof(*m));
  const PetscInt nnz = i[n];
  m->n = n; m->ncols = n; m->refct = 1; snprintf(m->type, sizeof m->type, "seqaij");
  m->i = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n + 1));
  m->j = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(nnz > 0 ? nnz : 1));
  m->a = (PetscScalar *)malloc(sizeof(PetscScalar) * (size_t)(nnz > 0 ? nnz : 1));
  memcpy(m->i, i, sizeof(PetscInt) * (size_t)(n + 1));
  memcpy(m->j, j, sizeof(PetscInt) * (size_t)nnz);
  memcpy(m->a, a, sizeof(PetscScalar) * (size_t)nnz);
  *A = m; return 0;
}
PetscErrorCode MatDestroy(Mat *A) {
  if (!A || !*A) return 0;
  if (--(*A)->refct <= 0) {
    if ((*A)->ops->destroy) (*A)->ops->destroy(*A);
    if ((*A)->i || (*A)->j) { free((*A)->i); free((*A)->j); free((*A)->a); }   /* (SeqDense borrows its array) */
    free(*A);
  }
  *A = NULL; return 0;
}
PetscErrorCode MatCreateSeqDense(PetscInt n, PetscInt ncols, PetscScalar *data, Mat *A) {
  Mat m = (Mat)calloc(1, sizeof(*m));
  m->n = n; m->ncols = ncols; m->a = data; m->refct = 1; snprintf(m->type, sizeof m->type, "seqdense");
  *A = m; return 0;
}
static struct { char name[16]; MatCreateFn fn; } g_mat[8];
static int g_nmat;
PetscErrorCode MatRegister(const char *type, MatCreateFn fn) {
  for (int i = 0; i < g_nmat; ++i) if (!strcmp(g_mat[i].name, type)) { g_mat[i].fn = fn; return 0; }
  if (g_nmat >= 8) SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "too many Mat types");
  snprintf(g_mat[g_nmat].name, 16, "%s", type); g_mat[g_nmat].fn = fn; ++g_nmat; return 0;
}
PetscErrorCode MatCreate(Mat *A) { Mat m = (Mat)calloc(1, sizeof(*m)); m->refct = 1; *A = m; return 0; }
PetscErrorCode MatSetType(Mat A, const char *type) {
  for (int i = 0; i < g_nmat; ++i) if (!strcmp(g_mat[i].name, type)) { snprintf(A->type, sizeof A->type, "%s", type); return g_mat[i].fn(A); }
  SETERRQ(PETSC_ERR_ARG_OUTOFRANGE, "Unknown Mat type %s", type);
}
#define MAT_OP(A, op, ...) do { if (!(A)->ops->op) SETERRQ(PETSC_ERR_SUP, "Mat type %s has no " #op, (A)->type); return (A)->ops->op(__VA_ARGS__); } while (0)
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { MAT_OP(A, mult, A, x, y); }
PetscErrorCode MatLUFactor(Mat A, IS row, IS col, const void *info) { MAT_OP(A, lufactor, A, row, col, info); }
PetscErrorCode MatSolve(Mat A, Vec b, Vec x) { MAT_OP(A, solve, A, b, x); }
PetscErrorCode MatMatSolve(Mat A, Mat B, Mat X) { MAT_OP(A, matsolve, A, B, X); }
PetscErrorCode MatGetDiagonal(Mat A, Vec d) { MAT_OP(A, getdiagonal, A, d); }
PetscErrorCode MatView(Mat A, char *buf, size_t len) { if (len) buf[0] = 0; MAT_OP(A, view, A, buf, len); }
PetscErrorCode VecCreateSeqWithArray(PetscInt n, PetscScalar *a, Vec *v) { Vec x = (Vec)calloc(1, sizeof(*x)); x->n = n; x->a = a; *v = x; return 0; }
PetscErrorCode VecDestroy(Vec *v) { if (v && *v) { free(*v); *v = NULL; } return 0; }
PetscErrorCode ISCreateGeneral(PetscInt n, const PetscInt *idx, IS *is) {
  IS s = (IS)calloc(1, sizeof(*s)); s->n = n; s->idx = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  memcpy(s->idx, idx, sizeof(PetscInt) * (size_t)n); *is = s; return 0; }
PetscErrorCode ISDestroy(IS *is) { if (is && *is) { free((*is)->idx); free(*is); *is = NULL; } return 0; }

PetscErrorCode PCCreate(PC *pc) { *pc = (PC)calloc(1, sizeof(**pc)); return 0; }
PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P) { pc->mat = A; pc->pmat = P; pc->setupcalled = 0; return 0; }
PetscErrorCode PCSetUp(PC pc) {
  if (!pc->ops->setup) SETERRQ(PETSC_ERR_ARG_WRONGSTATE, "PC has no type");
  PetscErrorCode ierr = pc->ops->setup(pc); CHKERRQ(ierr);
  pc->setupcalled = 1; return 0;
}
PetscErrorCode PCApply(PC pc, Vec x, Vec y) {
  if (!pc->setupcalled) { PetscErrorCode ierr = PCSetUp(pc); CHKERRQ(ierr); }
  return pc->ops->apply(pc, x, y);
}
PetscErrorCode PCDestroy(PC *pc) {
  if (!pc || !*pc) return 0;
  if ((*pc)->ops->destroy) (*pc)->ops->destroy(*pc);
  free(*pc); *pc = NULL; return 0;
}
PetscErrorCode KSPCreate(KSP *ksp) { *ksp = (KSP)calloc(1, sizeof(**ksp)); (*ksp)->rtol = 1e-5; (*ksp)->max_it = 10000; return 0; }
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat M) { ksp->A = A; ksp->M = M; ksp->setupstage = 0; return 0; }
/* KSPSolve sets up once per operator (PETSc's setupstage): KSPSetOperators / KSPSetFromOptions ask for a new setup,
 * repeated solves with the same operator reuse orderings, band and factors -- the reference's guard at
 * src/matbanded.c:171 relies on exactly that. */
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x) {
  PetscErrorCode ierr;
  ksp->vec_rhs = b; ksp->vec_sol = x;
  if (!ksp->ops->solve) SETERRQ(PETSC_ERR_ARG_WRONGSTATE, "KSP has no type");
  if (!ksp->setupstage) { ierr = ksp->ops->setup(ksp); CHKERRQ(ierr); ksp->setupstage = 1; }
  return ksp->ops->solve(ksp);
}
PetscErrorCode KSPDestroy(KSP *ksp) {
  if (!ksp || !*ksp) return 0;
  if ((*ksp)->ops->destroy) (*ksp)->ops->destroy(*ksp);
  free(*ksp); *ksp = NULL; return 0;
}

PetscErrorCode KSPSetOptionsPrefix(KSP ksp, const char *prefix) { snprintf(ksp->prefix, sizeof ksp->prefix, "%s", prefix ? prefix : ""); return 0; }
PetscErrorCode KSPSetFromOptions(KSP ksp) { ksp->setupstage = 0; return ksp->ops->setfromoptions ? ksp->ops->setfromoptions(ksp) : 0; }
PetscErrorCode KSPView(KSP ksp, char *buf, size_t len) { if (len) buf[0] = 0; return ksp->ops->view ? ksp->ops->view(ksp, buf, len) : 0; }
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its) { *its = ksp->its; return 0; }
PetscErrorCode KSPGetConvergedReason(KSP ksp, int *reason) { *reason = ksp->reason; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm) { *rnorm = ksp->rnorm; return 0; }
PetscErrorCode PCSetOptionsPrefix(PC pc, const char *prefix) { snprintf(pc->prefix, sizeof pc->prefix, "%s", prefix ? prefix : ""); return 0; }
PetscErrorCode PCSetFromOptions(PC pc) { return pc->ops->setfromoptions ? pc->ops->setfromoptions(pc) : 0; }
PetscErrorCode PCView(PC pc, char *buf, size_t len) { if (len) buf[0] = 0; return pc->ops->view ? pc->ops->view(pc, buf, len) : 0; }
