/*
 * matio.c -- the on-disk formats of the reference's drivers, for running real matrices through the path:
 *   PETSc binary Mat / Vec   read by MatLoad / VecLoad in /root/reference/src/testbed2.c:93-96, src/testbed.c
 *   MatrixMarket ASCII       written by MatView(PETSC_VIEWER_ASCII_MATRIXMARKET) in /root/reference/src/wbm.c:520-523
 * PETSc binary layout [EXTERNAL: PETSc manual, MatLoad]: big-endian int32 header {MAT_FILE_CLASSID = 1211216,
 * rows, cols, nnz}, int32 row lengths[rows], int32 column indices[nnz], float64 values[nnz];
 * Vec: {VEC_FILE_CLASSID = 1211214, n}, float64 values[n].  32-bit PetscInt, real PetscScalar (the reference's
 * configuration).  All functions return 0 on success and a PETSc-style non-zero error code otherwise;
 * arrays returned through pointers are malloc'd and owned by the caller (free()).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SPK_MAT_FILE_CLASSID 1211216
#define SPK_VEC_FILE_CLASSID 1211214
#define SPK_ERR_FILE_OPEN 65       /* PETSC_ERR_FILE_OPEN */
#define SPK_ERR_FILE_READ 66       /* PETSC_ERR_FILE_READ */
#define SPK_ERR_FILE_WRITE 67      /* PETSC_ERR_FILE_WRITE */
#define SPK_ERR_FILE_UNEXPECTED 79 /* PETSC_ERR_FILE_UNEXPECTED */
#define SPK_ERR_MEM 55             /* PETSC_ERR_MEM */

static uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
static uint64_t bswap64(uint64_t v) { return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32)); }
static int host_is_little(void) { const uint16_t one = 1; return *(const uint8_t*)&one == 1; }

static int read_i32(FILE* f, int32_t* dst, size_t n) {
  if (fread(dst, 4, n, f) != n) return SPK_ERR_FILE_READ;
  if (host_is_little()) for (size_t i = 0; i < n; ++i) dst[i] = (int32_t)bswap32((uint32_t)dst[i]);
  return 0;
}
static int read_f64(FILE* f, double* dst, size_t n) {
  if (fread(dst, 8, n, f) != n) return SPK_ERR_FILE_READ;
  if (host_is_little()) for (size_t i = 0; i < n; ++i) { uint64_t u; memcpy(&u, dst + i, 8); u = bswap64(u); memcpy(dst + i, &u, 8); }
  return 0;
}
static int write_i32(FILE* f, const int32_t* src, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u = (uint32_t)src[i];
    if (host_is_little()) u = bswap32(u);
    if (fwrite(&u, 4, 1, f) != 1) return SPK_ERR_FILE_WRITE;
  }
  return 0;
}
static int write_f64(FILE* f, const double* src, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint64_t u; memcpy(&u, src + i, 8);
    if (host_is_little()) u = bswap64(u);
    if (fwrite(&u, 8, 1, f) != 1) return SPK_ERR_FILE_WRITE;
  }
  return 0;
}

/* MatLoad of a PETSc binary AIJ matrix into 0-based CSR (ia: rows+1 offsets). */
int SpkMatLoadBinary(const char* path, int* rows, int* cols, int** ia, int** ja, double** a) {
  FILE* f = fopen(path, "rb");
  if (!f) return SPK_ERR_FILE_OPEN;
  int32_t hdr[4];
  int rc = read_i32(f, hdr, 4);
  if (!rc && hdr[0] != SPK_MAT_FILE_CLASSID) rc = SPK_ERR_FILE_UNEXPECTED;
  if (!rc && (hdr[1] < 0 || hdr[2] < 0 || hdr[3] < 0)) rc = SPK_ERR_FILE_UNEXPECTED;   /* nnz = -1: dense format, not used by the drivers */
  int32_t *len = NULL, *j = NULL, *off = NULL;
  double* v = NULL;
  if (!rc) {
    const size_t m = (size_t)hdr[1], nz = (size_t)hdr[3];
    len = (int32_t*)malloc(4 * (m ? m : 1)); off = (int32_t*)malloc(4 * (m + 1));
    j = (int32_t*)malloc(4 * (nz ? nz : 1)); v = (double*)malloc(8 * (nz ? nz : 1));
    if (!len || !off || !j || !v) rc = SPK_ERR_MEM;
    if (!rc) rc = read_i32(f, len, m);
    if (!rc) {
      int64_t s = 0;
      off[0] = 0;
      for (size_t i = 0; i < m; ++i) { if (len[i] < 0) { rc = SPK_ERR_FILE_UNEXPECTED; break; } s += len[i]; off[i + 1] = (int32_t)s; }
      if (!rc && s != (int64_t)nz) rc = SPK_ERR_FILE_UNEXPECTED;
    }
    if (!rc) rc = read_i32(f, j, nz);
    if (!rc) rc = read_f64(f, v, nz);
    if (!rc) for (size_t e = 0; e < nz; ++e) if (j[e] < 0 || j[e] >= hdr[2]) { rc = SPK_ERR_FILE_UNEXPECTED; break; }
  }
  fclose(f);
  free(len);
  if (rc) { free(off); free(j); free(v); return rc; }
  *rows = hdr[1]; *cols = hdr[2]; *ia = off; *ja = j; *a = v;
  return 0;
}

int SpkMatWriteBinary(const char* path, int rows, int cols, const int* ia, const int* ja, const double* a) {
  FILE* f = fopen(path, "wb");
  if (!f) return SPK_ERR_FILE_OPEN;
  const int32_t hdr[4] = {SPK_MAT_FILE_CLASSID, rows, cols, ia[rows]};
  int rc = write_i32(f, hdr, 4);
  for (int i = 0; !rc && i < rows; ++i) { const int32_t l = ia[i + 1] - ia[i]; rc = write_i32(f, &l, 1); }
  if (!rc) rc = write_i32(f, ja, (size_t)ia[rows]);
  if (!rc) rc = write_f64(f, a, (size_t)ia[rows]);
  if (fclose(f) && !rc) rc = SPK_ERR_FILE_WRITE;
  return rc;
}

int SpkVecLoadBinary(const char* path, int* n, double** v) {
  FILE* f = fopen(path, "rb");
  if (!f) return SPK_ERR_FILE_OPEN;
  int32_t hdr[2];
  int rc = read_i32(f, hdr, 2);
  if (!rc && (hdr[0] != SPK_VEC_FILE_CLASSID || hdr[1] < 0)) rc = SPK_ERR_FILE_UNEXPECTED;
  double* x = NULL;
  if (!rc) {
    x = (double*)malloc(8 * (hdr[1] ? (size_t)hdr[1] : 1));
    if (!x) rc = SPK_ERR_MEM;
    if (!rc) rc = read_f64(f, x, (size_t)hdr[1]);
  }
  fclose(f);
  if (rc) { free(x); return rc; }
  *n = hdr[1]; *v = x;
  return 0;
}

int SpkVecWriteBinary(const char* path, int n, const double* v) {
  FILE* f = fopen(path, "wb");
  if (!f) return SPK_ERR_FILE_OPEN;
  const int32_t hdr[2] = {SPK_VEC_FILE_CLASSID, n};
  int rc = write_i32(f, hdr, 2);
  if (!rc) rc = write_f64(f, v, (size_t)n);
  if (fclose(f) && !rc) rc = SPK_ERR_FILE_WRITE;
  return rc;
}

/* MatView(..., PETSC_VIEWER_ASCII_MATRIXMARKET): "coordinate real general", 1-based, row-major entry order.
 * PETSc prints values with %g; `digits` > 0 selects %.<digits>g instead (17 = exact round trip). */
int SpkMatWriteMatrixMarket(const char* path, int rows, int cols, const int* ia, const int* ja, const double* a, int digits) {
  FILE* f = fopen(path, "w");
  if (!f) return SPK_ERR_FILE_OPEN;
  int rc = 0;
  if (fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n", rows, cols, ia[rows]) < 0) rc = SPK_ERR_FILE_WRITE;
  for (int i = 0; !rc && i < rows; ++i)
    for (int e = ia[i]; e < ia[i + 1]; ++e) {
      const int w = digits > 0 ? fprintf(f, "%d %d %.*g\n", i + 1, ja[e] + 1, digits, a[e]) : fprintf(f, "%d %d %g\n", i + 1, ja[e] + 1, a[e]);
      if (w < 0) { rc = SPK_ERR_FILE_WRITE; break; }
    }
  if (fclose(f) && !rc) rc = SPK_ERR_FILE_WRITE;
  return rc;
}

void SpkFree(void* p) { free(p); }
