/*
 * ordering.c -- host orderings that feed the GPU gathers: reverse Cuthill-McKee, and the deterministic stand-in for
 * the reference's "fiedler" ordering.
 *
 * /root/reference/src/petsc_mat_fiedler.c:11-58 calls the proprietary HSL MC73 (multilevel Fiedler ordering,
 * src/hslmc73f.F90:16-30), which is not in the reference tree (src/makefile:6 links -lhsl_mc73) and cannot be
 * reproduced.  Its ROLE on the path -- a symmetric bandwidth-reducing permutation, returned as
 * row = col = "old index at every new position" (:49,54-56) -- is kept; the algorithm is replaced by reverse
 * Cuthill-McKee on the symmetrised pattern (George-Liu pseudo-peripheral start node, neighbours by ascending degree,
 * ties by index: fully deterministic), the same recipe the reference's own drivers use through PETSc's "rcm"
 * (src/HOWTO:2) and src/spectralPartition.c:379-388 applies per half.  Orderings are host setup per north_star; the
 * resulting IS is an INPUT shared by the CPU oracle and the GPU path, which applies it with gather kernels.
 * Registered like the reference's orderings (src/testbed2.c:66-68): MatOrderingRegister("fiedler", MatGetOrdering_Fiedler).
 */
#include "petsc_access.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* adjacency of the pattern of A + A^T without the diagonal: xadj[n+1], adj[] (columns ascending per row) */
static int sym_pattern(PetscInt n, const PetscInt *ai, const PetscInt *aj, PetscInt **xadj_o, PetscInt **adj_o) {
  PetscInt *cnt = (PetscInt *)calloc((size_t)n + 1, sizeof(PetscInt));
  if (!cnt) return 1;
  for (PetscInt r = 0; r < n; ++r) for (PetscInt q = ai[r]; q < ai[r + 1]; ++q) { const PetscInt c = aj[q]; if (c != r) { cnt[r + 1]++; cnt[c + 1]++; } }
  for (PetscInt r = 0; r < n; ++r) cnt[r + 1] += cnt[r];
  const PetscInt tot = cnt[n];
  PetscInt *tmp = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(tot > 0 ? tot : 1));
  PetscInt *pos = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n + 1));
  if (!tmp || !pos) { free(cnt); free(tmp); free(pos); return 1; }
  memcpy(pos, cnt, sizeof(PetscInt) * (size_t)(n + 1));
  /* rows are visited in ascending order, so every list comes out sorted ascending (with duplicates) */
  for (PetscInt r = 0; r < n; ++r) for (PetscInt q = ai[r]; q < ai[r + 1]; ++q) { const PetscInt c = aj[q]; if (c != r) { tmp[pos[r]++] = c; tmp[pos[c]++] = r; } }
  /* the two insert streams of a row (its own columns, the rows that reference it) interleave: sort + unique */
  PetscInt *xadj = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n + 1));
  PetscInt w = 0; xadj[0] = 0;
  for (PetscInt r = 0; r < n; ++r) {
    PetscInt *lst = tmp + cnt[r]; const PetscInt len = cnt[r + 1] - cnt[r];
    for (PetscInt a = 1; a < len; ++a) { const PetscInt v = lst[a]; PetscInt b = a - 1; while (b >= 0 && lst[b] > v) { lst[b + 1] = lst[b]; --b; } lst[b + 1] = v; }
    PetscInt last = -1;
    for (PetscInt a = 0; a < len; ++a) if (lst[a] != last) { tmp[w++] = lst[a]; last = lst[a]; }
    xadj[r + 1] = w;
  }
  free(cnt); free(pos);
  *xadj_o = xadj; *adj_o = tmp;
  return 0;
}

/* BFS from `root` over unvisited nodes (mask[v] == 0); fills order[lo..) level by level, neighbours by (degree, index);
 * returns the count, *nlev = number of levels, *last_lo = start of the last level inside order */
static PetscInt bfs_levels(PetscInt root, const PetscInt *xadj, const PetscInt *adj, PetscInt *mask, PetscInt *order, PetscInt lo,
                           PetscInt *nlev, PetscInt *last_lo, PetscInt *scratch) {
  PetscInt head = lo, tail = lo, lev = 0, lev_end;
  order[tail++] = root; mask[root] = 1;
  lev_end = tail; *last_lo = lo;
  while (head < tail) {
    const PetscInt v = order[head++];
    PetscInt m = 0;
    for (PetscInt q = xadj[v]; q < xadj[v + 1]; ++q) { const PetscInt u = adj[q]; if (!mask[u]) { mask[u] = 1; scratch[m++] = u; } }
    /* ascending degree, ties by index (adj is index-sorted, insertion sort is stable) */
    for (PetscInt a = 1; a < m; ++a) {
      const PetscInt u = scratch[a], du = xadj[u + 1] - xadj[u]; PetscInt b = a - 1;
      while (b >= 0 && (xadj[scratch[b] + 1] - xadj[scratch[b]]) > du) { scratch[b + 1] = scratch[b]; --b; }
      scratch[b + 1] = u;
    }
    for (PetscInt a = 0; a < m; ++a) order[tail++] = scratch[a];
    if (head == lev_end && tail > lev_end) { ++lev; *last_lo = lev_end; lev_end = tail; }
  }
  *nlev = lev + 1;
  return tail - lo;
}

/* perm[new] = old (what MatPermute / the GPU gathers take).  0 on success. */
int SpkOrderingRCM(PetscInt n, const PetscInt *ai, const PetscInt *aj, PetscInt *perm) {
  PetscInt *xadj, *adj;
  if (sym_pattern(n, ai, aj, &xadj, &adj)) return 1;
  PetscInt *mask = (PetscInt *)calloc((size_t)(n > 0 ? n : 1), sizeof(PetscInt));
  PetscInt *order = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  PetscInt *scratch = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  if (!mask || !order || !scratch) { free(xadj); free(adj); free(mask); free(order); free(scratch); return 1; }
  PetscInt done = 0;
  for (PetscInt s = 0; s < n; ++s) {
    if (mask[s]) continue;
    /* pseudo-peripheral node of this component (George & Liu): repeat BFS from the minimum-degree node of the
     * last level while the eccentricity grows */
    PetscInt root = s, nlev = 0, last_lo = 0, cnt;
    for (int iter = 0; iter < 12; ++iter) {
      cnt = bfs_levels(root, xadj, adj, mask, order, done, &nlev, &last_lo, scratch);
      PetscInt best = order[last_lo], bd = xadj[best + 1] - xadj[best];
      for (PetscInt q = last_lo; q < done + cnt; ++q) { const PetscInt u = order[q], du = xadj[u + 1] - xadj[u]; if (du < bd || (du == bd && u < best)) { best = u; bd = du; } }
      for (PetscInt q = done; q < done + cnt; ++q) mask[order[q]] = 0;
      PetscInt nlev2, ll2;
      const PetscInt cnt2 = bfs_levels(best, xadj, adj, mask, order, done, &nlev2, &ll2, scratch);
      for (PetscInt q = done; q < done + cnt2; ++q) mask[order[q]] = 0;
      if (nlev2 <= nlev || best == root) break;
      root = best;
    }
    cnt = bfs_levels(root, xadj, adj, mask, order, done, &nlev, &last_lo, scratch);   /* Cuthill-McKee order of the component */
    done += cnt;
  }
  for (PetscInt i = 0; i < n; ++i) perm[i] = order[n - 1 - i];                        /* reversed */
  free(xadj); free(adj); free(mask); free(order); free(scratch);
  return 0;
}

PetscErrorCode MatGetOrdering_RCM(Mat A, const char *type, IS *row, IS *col) {
  PetscInt n; const PetscInt *ai, *aj; const PetscScalar *aa; PetscErrorCode ierr;
  (void)type;
  ierr = SpkMatGetCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  PetscInt *perm = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  if (!perm || SpkOrderingRCM(n, ai, aj, perm)) { free(perm); SPK_ERR(PETSC_ERR_LIB, "RCM ordering: out of memory"); }
  ierr = SpkMatRestoreCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  ierr = ISCreateGeneral(n, perm, row);CHKERRQ(ierr);     /* row = col: a symmetric permutation (src/petsc_mat_fiedler.c:54-56) */
  ierr = ISCreateGeneral(n, perm, col);CHKERRQ(ierr);
  free(perm);
  return 0;
}
/* the "fiedler" slot of the reference (MC73 absent): see the header of this file */
PetscErrorCode MatGetOrdering_Fiedler(Mat A, const char *type, IS *row, IS *col) { return MatGetOrdering_RCM(A, type, row, col); }

/* "awbm" (src/petsc_mat_awbm.c:42-225, registered at src/testbed2.c:67) with the matching computed on the GPU by
 * spk_awbm_csr: row IS = permR, column IS = identity (:201-205); the scalings are computed and dropped like in the
 * reference (:221-222).  A short-lived engine context carries the device and the error text. */
#include "../../include/spike_b200.h"
PetscErrorCode MatGetOrdering_AWBM(Mat A, const char *type, IS *row, IS *col) {
  PetscInt n; const PetscInt *ai, *aj; const PetscScalar *aa; PetscErrorCode ierr;
  (void)type;
  ierr = SpkMatGetCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  spk_ctx *ctx = NULL; spk_opts o;
  spk_default_opts(&o);
  if (spk_create(&ctx, &o)) SPK_ERR(PETSC_ERR_LIB, "AWBM ordering: %s", spk_last_error(NULL));
  PetscInt *pr = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  PetscInt *pc = (PetscInt *)malloc(sizeof(PetscInt) * (size_t)(n > 0 ? n : 1));
  if (!pr || !pc) { free(pr); free(pc); spk_destroy(&ctx); SPK_ERR(PETSC_ERR_LIB, "AWBM ordering: out of memory"); }
  const int rc = spk_awbm_csr(ctx, (int)n, (const int *)ai, (const int *)aj, aa, (int *)pr, NULL, NULL, NULL, NULL);
  if (rc) {
    char msg[256]; snprintf(msg, sizeof msg, "%s", spk_last_error(ctx));
    free(pr); free(pc); spk_destroy(&ctx);
    SPK_ERR(PETSC_ERR_LIB, "AWBM ordering failed (%d): %s", rc, msg);
  }
  spk_destroy(&ctx);
  ierr = SpkMatRestoreCSR(A, &n, &ai, &aj, &aa);CHKERRQ(ierr);
  for (PetscInt i = 0; i < n; ++i) pc[i] = i;
  ierr = ISCreateGeneral(n, pr, row);CHKERRQ(ierr);
  ierr = ISCreateGeneral(n, pc, col);CHKERRQ(ierr);
  free(pr); free(pc);
  return 0;
}
