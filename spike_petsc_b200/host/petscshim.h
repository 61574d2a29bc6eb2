/*
 * petscshim.h -- the few PETSc-shaped types the host glue needs, so that the C mirror of the
 * reference's plugin surface (PCBANDED, KSPREORDER, MatCreateSubMatrixBanded, MatOrdering hooks)
 * compiles and runs in an image without PETSc.  With -DHAVE_PETSC the glue includes the real
 * headers instead and these definitions vanish (INTEGRATION.md shows the registration calls).
 *
 * Shapes follow the reference: per-type ops tables filled at create time
 * (/root/reference/src/matbanded.c:264-273, src/kspreorder.c:210-216), private state behind
 * ->data, int error codes with 0 = success, options looked up by prefixed name.
 */
#ifndef SPK_PETSCSHIM_H
#define SPK_PETSCSHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int    PetscErrorCode;
typedef int    PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_SUP 56
#define PETSC_ERR_LIB 76

/* SeqAIJ-like matrix: 0-based CSR, rows sorted by column (what MatGetRow returns).  Other Mat types (SeqDense:
 * a = column-major n x ncols, i = j = NULL; MATBANDED: everything behind ->data) fill the ops table, as PETSc's
 * MatCreate_XXX constructors do. */
typedef struct _p_Mat *Mat;
typedef struct _p_Vec { PetscInt n; PetscScalar *a; } *Vec;
typedef struct _p_IS  { PetscInt n; PetscInt *idx; } *IS;
struct _MatOps {
  PetscErrorCode (*mult)(Mat, Vec, Vec);
  PetscErrorCode (*lufactor)(Mat, IS, IS, const void *info);
  PetscErrorCode (*solve)(Mat, Vec, Vec);
  PetscErrorCode (*matsolve)(Mat, Mat, Mat);
  PetscErrorCode (*getdiagonal)(Mat, Vec);
  PetscErrorCode (*view)(Mat, char *buf, size_t len);
  PetscErrorCode (*destroy)(Mat);
};
struct _p_Mat { PetscInt n; PetscInt *i, *j; PetscScalar *a; int refct; PetscInt ncols; struct _MatOps ops[1]; void *data; char type[16]; };

typedef struct _p_PC *PC;
struct _PCOps {
  PetscErrorCode (*setup)(PC);
  PetscErrorCode (*apply)(PC, Vec, Vec);
  PetscErrorCode (*applytranspose)(PC, Vec, Vec);
  PetscErrorCode (*reset)(PC);
  PetscErrorCode (*destroy)(PC);
  PetscErrorCode (*setfromoptions)(PC);
  PetscErrorCode (*view)(PC, char *buf, size_t len);
};
struct _p_PC { struct _PCOps ops[1]; Mat mat, pmat; int setupcalled; void *data; char prefix[160]; };

typedef struct _p_KSP *KSP;
struct _KSPOps {
  PetscErrorCode (*setup)(KSP);
  PetscErrorCode (*solve)(KSP);
  PetscErrorCode (*destroy)(KSP);
  PetscErrorCode (*setfromoptions)(KSP);
  PetscErrorCode (*view)(KSP, char *buf, size_t len);
};
typedef PetscErrorCode (*MatOrderingFn)(Mat, const char *type, IS *row, IS *col);
struct _p_KSP {
  int setupstage;      /* 0: KSPSetUp needed (new operators / options), 1: set up */
  struct _KSPOps ops[1];
  Mat A, M;            /* operators (KSPSetOperators) */
  Vec vec_rhs, vec_sol;
  PC  pc;
  int reason, its; PetscReal rnorm, rtol; PetscInt max_it;
  void *data; char prefix[128];
};

/* options database: "-name value" pairs, looked up with the object's prefix */
PetscErrorCode PetscOptionsSetValue(const char *name, const char *value);
PetscErrorCode PetscOptionsClear(void);
PetscErrorCode PetscOptionsGetInt(const char *prefix, const char *name, PetscInt *v, PetscBool *set);
PetscErrorCode PetscOptionsGetReal(const char *prefix, const char *name, PetscReal *v, PetscBool *set);
PetscErrorCode PetscOptionsGetString(const char *prefix, const char *name, char *v, size_t len, PetscBool *set);

/* ordering registry: MatOrderingRegister(name, fn) as in src/testbed2.c:66-68 */
PetscErrorCode MatOrderingRegister(const char *name, MatOrderingFn fn);
PetscErrorCode MatGetOrdering(Mat, const char *type, IS *row, IS *col);

/* object helpers */
PetscErrorCode MatCreateSeqAIJWithArrays(PetscInt n, const PetscInt *i, const PetscInt *j, const PetscScalar *a, Mat *A);
PetscErrorCode MatDestroy(Mat *A);
PetscErrorCode MatCreateSeqDense(PetscInt n, PetscInt ncols, PetscScalar *data, Mat *A);   /* column-major, data borrowed */
typedef PetscErrorCode (*MatCreateFn)(Mat);
PetscErrorCode MatRegister(const char *type, MatCreateFn fn);     /* MatRegister("banded", MatCreate_Banded) */
PetscErrorCode MatCreate(Mat *A);
PetscErrorCode MatSetType(Mat A, const char *type);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatLUFactor(Mat A, IS row, IS col, const void *info);
PetscErrorCode MatSolve(Mat A, Vec b, Vec x);
PetscErrorCode MatMatSolve(Mat A, Mat B, Mat X);
PetscErrorCode MatGetDiagonal(Mat A, Vec d);
PetscErrorCode MatView(Mat A, char *buf, size_t len);
PetscErrorCode VecCreateSeqWithArray(PetscInt n, PetscScalar *a, Vec *v);
PetscErrorCode VecDestroy(Vec *v);
PetscErrorCode ISCreateGeneral(PetscInt n, const PetscInt *idx, IS *is);
PetscErrorCode ISDestroy(IS *is);
PetscErrorCode PCCreate(PC *pc);
PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P);
PetscErrorCode PCSetUp(PC pc);
PetscErrorCode PCApply(PC pc, Vec x, Vec y);
PetscErrorCode PCDestroy(PC *pc);
PetscErrorCode KSPCreate(KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat M);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPDestroy(KSP *ksp);
PetscErrorCode KSPSetOptionsPrefix(KSP ksp, const char *prefix);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPView(KSP ksp, char *buf, size_t len);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetConvergedReason(KSP ksp, int *reason);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm);
PetscErrorCode PCSetOptionsPrefix(PC pc, const char *prefix);
PetscErrorCode PCSetFromOptions(PC pc);
PetscErrorCode PCView(PC pc, char *buf, size_t len);
const char *PetscLastErrorMessage(void);
void PetscSetErrorMessage(const char *fmt, ...);

#define SETERRQ(code, ...) do { PetscSetErrorMessage(__VA_ARGS__); return (code); } while (0)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)

#ifdef __cplusplus
}
#endif
#endif
