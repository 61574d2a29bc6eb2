/* Minimal stand-in for <petscsys.h> so that /root/reference/src/hslmc64.c (f2c MC64) compiles
 * stand-alone into oracle/_ref/ (see oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
 * Only the typedefs/macros that file uses (symbol census in SURVEY.md 8c). */
#ifndef ORACLE_SHIM_PETSCSYS_H
#define ORACLE_SHIM_PETSCSYS_H
#include <math.h>
#include <stdlib.h>
typedef int    PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int    PetscErrorCode;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
#ifdef __cplusplus
#define PETSC_EXTERN extern "C"
#else
#define PETSC_EXTERN extern
#endif
#define PetscFunctionBegin
#define PetscFunctionReturn(a) return (a)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscAbsScalar(a) fabs(a)
#define PETSC_COMM_SELF 0
#define PETSC_ERR_SUP 56
#define SETERRQ(comm, code, msg) return (code)
#endif
