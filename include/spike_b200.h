/*
 * spike_b200.h -- C ABI of the B200-native SPIKE banded engine (libspike_b200.so).
 *
 * This is the drop-in boundary for the hot path of spikegpu/spike-petsc: plain C, opaque handle,
 * pointers + sizes, int return codes (0 = success, never throws/aborts -- the PetscErrorCode
 * convention of the reference, e.g. src/matbanded.c:95, src/kspreorder.c:121).  Each entry point
 * names the reference interface it replaces (paths relative to the reference tree).  The
 * PETSc-shaped glue that binds these calls into PC/KSP ops tables is spike_petsc_b200/host/
 * (see INTEGRATION.md).
 *
 * Threading: one host thread per context (PETSc objects are not thread safe).
 * All work is enqueued on the context's CUDA stream; calls that return host-visible results
 * synchronise that stream before returning.  One exception, invisible to the caller: the spike tips and reduced
 * blocks of a narrow-band spk_factor run on a context-owned side stream (ordered behind the band LU by an event) so
 * that the first spk_solve's partition sweeps need not wait for them; the reduced solve of that spk_solve -- and
 * every other entry point that reads the tips or rewrites the band -- makes the context stream wait for them first.
 * Work the caller enqueues on the context stream after an spk_solve / spk_view is therefore ordered after all of
 * the factorisation; SPIKE_B200_SIDE_STREAM=0 in the environment keeps everything on the one stream.  There is NO CPU fallback: every call fails with
 * SPK_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SPIKE_B200_H
#define SPIKE_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct spk_ctx spk_ctx;

enum { SPK_OK = 0, SPK_ERR_ARG = 1, SPK_ERR_CUDA = 2, SPK_ERR_STATE = 3, SPK_ERR_UNSUPPORTED = 4, SPK_ERR_NOMEM = 5 };
enum { SPK_MEM_HOST = 0, SPK_MEM_DEVICE = 1 };                    /* where caller buffers live      */
enum { SPK_LAYOUT_ROWS = 0, SPK_LAYOUT_DIAGS = 1 };               /* dense band input layouts       */
enum { SPK_KSP_GMRES = 0, SPK_KSP_BCGS = 1 };                     /* inner KSP types used by HOWTO  */

typedef struct spk_opts {
  int     device;        /* CUDA device ordinal                                                      */
  void   *stream;        /* cudaStream_t to enqueue on; NULL = legacy default stream                 */
  int     partitions;    /* SPIKE partitions on this device; 0 = auto (multiple of the SM count)     */
  int     tip_tiles;     /* truncation window, in 8-row tiles, of the W^(t) tip pass and of the apply
                            corrections; 0 = auto, <0 = whole partition (SaP-style full passes)      */
  double  boost_rel;     /* pivot boosting: |pivot| < boost_rel*max|a_ij| is replaced by +-that      */
  int     mem;           /* SPK_MEM_HOST / SPK_MEM_DEVICE for vector arguments of solve/mult/krylov  */
  int     rank, nranks;  /* this device's position in a row-block sharding (0,1 = single device)     */
  int64_t row_offset;    /* global index of this shard's first row (generator counters, bookkeeping) */
  int64_t n_global;      /* global rows; 0 = same as local n                                         */
} spk_opts;

typedef struct spk_info {
  int64_t n, n_padded;
  int     k, k_padded, kt;           /* half-bandwidth, padded to 8*kt                                 */
  int     partitions, tip_tiles;
  int64_t boosted_pivots;
  int     factored;
  double  frac;                      /* norm fraction of the extracted band (PC_Banded.f)             */
  double  anorm_max;                 /* max |a_ij| of the band                                        */
  double  factor_ms, solve_ms;       /* device time of the last spk_factor / spk_solve (CUDA events; 0 unless spk_set_timing(ctx,1)) */
  int64_t band_bytes;                /* bytes of the device band (tile-major, padded)                 */
  int     kernel_launches;           /* kernels launched by the last spk_factor + spk_solve           */
  /* device time (ms, CUDA events on the context stream) of the kernels of the last factor / solve:
     [0] bottom-up tip windows  [1] band LU  [2] spike tips + reduced factor
     [3] partition sweeps       [4] reduced solve  [5] corrections   [6],[7] reserved */
  double  stage_ms[8];
} spk_info;

void spk_default_opts(spk_opts *o);
const char *spk_last_error(const spk_ctx *ctx);       /* message of the last failing call            */
const char *spk_version(void);

/* lifecycle -- PCCreate_Banded / PCReset_Banded / PCDestroy_Banded (src/matbanded.c:251-283,120-145) */
int spk_create(spk_ctx **ctx, const spk_opts *opts);
int spk_destroy(spk_ctx **ctx);

/* ---- band definition -------------------------------------------------------------------------- */
/* Dense band in host or device memory.  ROWS: a[i*(2k+1) + (j-i+k)];  DIAGS: a[(j-i+k)*n + i].
 * Replaces holding B as AIJ (PC_Banded.B, src/matbanded.c:114).  For a shard (opts.nranks > 1) n is the number of
 * local rows and j the local column index: entries with j < 0 or j >= n that exist in the global matrix
 * (opts.row_offset, opts.n_global) are the coupling blocks to the neighbour ranks and are kept. */
int spk_set_band_dense(spk_ctx *ctx, int64_t n, int k, const double *band, int layout, int mem);

/* MatCreateSubMatrixBanded (src/matbanded.h:5, src/matbanded.c:22-107) fused with the MatPermute of
 * KSPSetUp_Reorder (src/kspreorder.c:20): B = band_k( A(rowperm[i], colperm[j]) ).  CSR is 0-based,
 * host memory; rowperm/colperm may be NULL (identity).  *kmax in/out and *frac in/out exactly as
 * the reference (k chosen in row-major summation order on the host, including the fall-through
 * quirk); the pack + gather runs on the GPU. */
int spk_set_band_csr(spk_ctx *ctx, int n, const int *ia, const int *ja, const double *a,
                     const int *rowperm, const int *colperm, int *kmax, double *frac);

/* Counter-based synthetic band generated directly in device memory:
 * a_ij = 2*u01(splitmix64(seed ^ (gi*(2k+1)+d+k))) - 1, a_ii = delta * sum|a_ij|, gi = global row. */
int spk_set_band_synthetic(spk_ctx *ctx, int64_t n, int k, uint64_t seed, double delta);

/* copy the device band (original or factored) back in ROWS layout -- test / debug hook */
int spk_get_band_rows(spk_ctx *ctx, double *band_rows_host);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* PCSetUp(b->pc) at src/matbanded.c:178 (and MatLUFactor of a MATBANDED): per-partition banded LU
 * (no pivoting, diagonal boosting), spike tips V^(b)/W^(t), reduced-system factorisation. In place.
 * Half-bandwidths up to 128 run the register-resident window kernel (csrc/lu.cu); 129..512 the super-block kernel
 * whose trailing updates are 64^3 FP64 tensor-core products (csrc/wide_lu.cu); wider bands are refused. */
int spk_factor(spk_ctx *ctx);
/* PCApply(b->pc,x,y) at src/matbanded.c:190 (and MatSolve / MatMatSolve): x = B^{-1} b for nrhs vectors of
 * leading dimension n.  b and x may alias.  nrhs >= 2 takes the block path: sweeps and window corrections run for
 * 8 columns per warp as tile products on the FP64 tensor cores and read the band once per 32 columns. */
int spk_solve(spk_ctx *ctx, const double *b, double *x, int nrhs);
/* MatMult (src/testbed2.c:122 and inside the inner KSP) with the UNFACTORED band kept by
 * spk_keep_original(ctx,1) or, before spk_factor, with the band itself. */
int spk_mult(spk_ctx *ctx, const double *x, double *y);
/* keep != 0: keep a copy of the unfactored band (call before the band is set, or before spk_factor).  spk_factor
 * then reads that copy and writes the factors into the working band (out of place, same traffic), so it can be
 * called again -- PCSetUp after PCReset with an unchanged operator -- without setting the band anew.  Without it
 * the factorisation is in place and a second spk_factor is an error. */
int spk_keep_original(spk_ctx *ctx, int keep);

/* VecPermute (src/kspreorder.c:122-127): inverse=0: v[i] <- v[idx[i]];  inverse=1: v[idx[i]] <- v[i].
 * idx is a host int array of length n; v follows opts.mem. */
int spk_permute(spk_ctx *ctx, const int *idx, int inverse, double *v, int64_t n);

/* Inner KSPSolve at src/kspreorder.c:124: left-preconditioned GMRES(restart) / BiCGStab on the
 * device with M^{-1} = this context's SPIKE solve.  A is the CSR matrix registered with
 * spk_set_operator_csr (permuted like the band) or, if none, the unfactored band. */
int spk_set_operator_csr(spk_ctx *ctx, int n, const int *ia, const int *ja, const double *a,
                         const int *rowperm, const int *colperm);
int spk_krylov(spk_ctx *ctx, int method, int restart, double rtol, int maxit, const double *b,
               double *x, int *its, double *rnorm, int *converged);

/* Equilibration: band <- diag(rscale) band diag(cscale) in place, after spk_set_band_* and before
 * spk_factor; spk_solve and the spk_krylov preconditioner then apply diag(c) (scaled band)^-1 diag(r), i.e. the
 * inverse of the ORIGINAL band.  rscale = exp(u), cscale = exp(v) of MC64 job 5 are the scalings the reference
 * computes and discards (src/petsc_mat_wbm.c:56; AWBM: src/petsc_mat_awbm.c:208-223).  n entries each, memory
 * space per opts.mem.  spk_mult keeps using the unscaled original.  Sharded contexts (nranks > 1) pass n + 2*kp
 * column scales: the left neighbour's last kp, their own n, the right neighbour's first kp (the coupling blocks
 * live in this rank's halo tiles); entries towards a missing neighbour are ignored. */
int spk_set_scaling(spk_ctx *ctx, const double *rscale, const double *cscale);

/* MatGetOrdering_AWBM (src/petsc_mat_awbm.c:42-225) on the GPU: approximate weighted bipartite
 * matching of a square CSR matrix, bit-identical to the reference's serial algorithm.  Weights, duals, edge tightness
 * and scalings are streaming kernels; the order-dependent greedy pass (:98-112) runs as rounds of proposals that
 * reproduce the serial outcome exactly; the repair passes for columns left without a free tight row (:115-193,
 * normally a handful of columns) finish on the host by the same serial rules.  All pointers are host memory.
 * permR[match[c]] = c is the row IS the reference returns (:201-203; its column IS is the identity, :204);
 * match (CSR row c -> matched index), scalR = exp(v)/amax and scalC = exp(u) (:212-215, computed and discarded by the
 * reference; usable with spk_set_scaling) and stats[4] = {device rounds, columns matched on the device, columns
 * finished by the serial greedy rule on the host, columns that needed a repair pass} are optional (NULL). */
int spk_awbm_csr(spk_ctx *ctx, int n, const int *ia, const int *ja, const double *a, int *permR, int *match,
                 double *scalR, double *scalC, int *stats);

/* Self-check of the factorisation (no reference counterpart: the reference's inner PC is an exact LU): with the
 * unfactored band kept (spk_keep_original), solves B x = B v for a fixed probe v and returns ||x - v|| / ||v||.  The
 * truncated SPIKE apply equals the exact band solve only where the spikes decay inside the truncation window
 * (diagonally dominant bands); elsewhere this is its error as an approximation of B^-1.  opts.partitions = 1 is the
 * exact mode (one partition, no truncation). */
int spk_check(spk_ctx *ctx, double *rel_err);

/* PCView_Banded (src/matbanded.c:196-211) */
int spk_view(spk_ctx *ctx, spk_info *info);
/* level 2: record the CUDA events behind spk_info.factor_ms / solve_ms / stage_ms in the following calls (the
 * per-stage timers + JSON records of SURVEY section 5; what -log_view would show of the reference's PCSetUp / PCApply);
 * level 1: only stage_ms[1], the band LU (one pair of records per factorisation); level 0: none.
 * 0 by default: every event record between two kernels costs about 2 us of stream time, 25-30 us per factor +
 * solve at level 2.  SPIKE_B200_STAGE_TIMERS=1 in the environment makes level 2 the default of every new context. */
int spk_set_timing(spk_ctx *ctx, int level);

/* ---- multi-GPU row-block sharding: spike-tip exchange hooks (the host moves these buffers between
 * neighbouring ranks with NCCL send/recv; sizes are (8*kt)^2 doubles for tips, 8*kt for vectors) -- */
int spk_tip_size(spk_ctx *ctx, int *kp);
int spk_get_boundary(spk_ctx *ctx, int which, double *dev_buf);   /* see SPK_BND_* */
int spk_set_boundary(spk_ctx *ctx, int which, const double *dev_buf);
enum {
  SPK_BND_WT_FIRST = 0,     /* out: W^(t) of my first partition (kp*kp)            -> left rank, once  */
  SPK_BND_REMOTE_WT = 2,    /* in : right neighbour's W^(t)                                           */
  SPK_BND_G_TOP = 3,        /* out: g^(t), first kp entries of my D^-1 b (per solve) -> left rank     */
  SPK_BND_REMOTE_G_TOP = 4, /* in : right neighbour's g^(t)                                           */
  SPK_BND_X_BOT = 5,        /* out: x^(b) of the boundary interface (per solve)     -> right rank     */
  SPK_BND_REMOTE_X_BOT = 6, /* in : left neighbour's x^(b)                                            */
  SPK_BND_HALO_LEFT = 7,    /* in : last kp entries of the left neighbour's x  (sharded spk_mult)     */
  SPK_BND_HALO_RIGHT = 8    /* in : first kp entries of the right neighbour's x (sharded spk_mult)    */
};
/* split-phase factor/solve for nranks > 1 (the host performs the NCCL exchanges between phases):
 *   factor: phase 0 (tip windows + LU), phase 1 (local tips), [WT_FIRST -> left's REMOTE_WT], phase 2
 *      or, overlapping the exchange with the LU: phase 10 (tip windows + every W^(t)), start
 *      [WT_FIRST -> left's REMOTE_WT], phase 11 (band LU), finish the exchange, phase 1 (local tips and,
 *      REMOTE_WT being set, the boundary block in the same launch), phase 2 (no-op then)
 *   solve : phase 0 (sweeps), [G_TOP -> left's REMOTE_G_TOP], phase 1 (reduced systems),
 *           [X_BOT -> right's REMOTE_X_BOT], phase 2 (corrections).  Device vectors.
 *           nrhs > 1 (column r at b + r*n): the vector items become kp x nrhs blocks (column r at + r*kp); reserve
 *           the exchange buffers once with spk_reserve_rhs(ctx, max nrhs) after the band is set and before
 *           spk_peer_mailbox_create.  Phases 1 and 2 use the nrhs of phase 0.
 * With nranks == 1 spk_factor / spk_solve run all phases back to back. */
int spk_factor_phase(spk_ctx *ctx, int phase);
int spk_solve_phase(spk_ctx *ctx, int phase, const double *b, double *x, int nrhs);
int spk_reserve_rhs(spk_ctx *ctx, int nrhs);

/* ---- the same exchanges through NVLink peer memory instead of host-driven NCCL send/recv (csrc/peer.cu).
 * Every rank owns a mailbox in device memory; neighbours map it with CUDA IPC (the 64-byte handle travels
 * once, over any host channel), a producer's kernel stores the item straight into the consumer's mailbox and
 * releases a sequence flag, the consumer's kernel acquires it, moves the item where spk_set_boundary would
 * have put it and acknowledges.  spk_peer_post(which) replaces spk_get_boundary(which) + send,
 * spk_peer_wait(which) replaces recv + spk_set_boundary(which); both only enqueue a kernel on the context's
 * stream (no host synchronisation), so the phase order above is unchanged and the W^(t) exchange overlaps the
 * band LU by construction.  Replaces, like the hooks above, the VecScatter / MatGetSubMatrices traffic PETSc
 * would issue for a distributed Mat on this path (src/matbanded.c:141-150 runs the reference on one rank only).
 *   spk_peer_mailbox_create : allocate my mailbox (after the band is set); handle64 <- cudaIpcMemHandle_t,
 *                             dev_ptr <- its device address (either may be NULL)
 *   spk_peer_mailbox_attach : side 0 = left neighbour, 1 = right; pass the neighbour's handle, or its device
 *                             address when it lives in the same process
 *   spk_peer_check          : synchronise the stream; SPK_ERR_STATE if a bounded spin expired
 * A spin that expires (SPIKE_B200_PEER_TIMEOUT_S seconds, default 20) poisons its destination with NaN, acknowledges
 * nothing, and makes every later factor / solve / post / wait call on the context fail until the band is set anew. */
int spk_peer_mailbox_create(spk_ctx *ctx, void *handle64, void **dev_ptr);
int spk_peer_mailbox_attach(spk_ctx *ctx, int side, const void *handle64, void *direct_ptr);
int spk_peer_post(spk_ctx *ctx, int which);   /* which: SPK_BND_WT_FIRST, SPK_BND_G_TOP, SPK_BND_X_BOT */
int spk_peer_wait(spk_ctx *ctx, int which);   /* which: SPK_BND_REMOTE_WT, SPK_BND_REMOTE_G_TOP, SPK_BND_REMOTE_X_BOT */
int spk_peer_check(spk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPIKE_B200_H */
