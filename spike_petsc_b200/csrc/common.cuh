// common.cuh -- shared definitions of the SPIKE engine (device band layout, context, helpers).
//
// Device band layout ("tile-major band"): the matrix is cut into 8x8 tiles; tile row I holds the
// 2*KT+1 tiles (I, I-KT .. I+KT) contiguously, each tile 64 doubles row-major (512 B):
//     tile(I,J) at  band + ((I*(2KT+1) + (J-I+KT)) * 64),   element (r,c) at  + r*8 + c
// KT = ceil(K/8).  Rows are padded to NT*8 with identity rows; tiles that stick out of the matrix
// are stored as zeros.  A tile is exactly one DMMA m8n8k4 accumulator fragment (lane l holds the
// 16 B at doubles 2l,2l+1), so LU tiles move between HBM and tensor-core registers with one
// coalesced 128-bit access per lane, and a tile row is one contiguous chunk for bulk copies.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdlib>
#include "../../include/spike_b200.h"

#define SPK_TILE 8
#define SPK_TILE_ELEMS 64
#define SPK_MAX_KT 16      // widest band (in tiles) of the register-resident kernels (lu.cu, solve.cu, msweep.cu, tips.cu)
#define SPK_MAX_K 512      // widest band overall: above 8*SPK_MAX_KT the super-block kernels of wide*.cu take over

struct BandLayout {
  int64_t n;    // rows of the matrix
  int64_t nt;   // tile rows (ceil(n/8))
  int k;        // half bandwidth
  int kt;       // ceil(k/8)
  int tpr;      // tiles per tile row = 2*kt+1
  int kc;       // tiles of a coupling block / spike tip (kp/8): = kt for narrow bands, 8*KB for wide ones (kt = 8*KB+7 there)
  __host__ __device__ int64_t tile_off(int64_t I, int64_t J) const { return (I * tpr + (J - I + kt)) * SPK_TILE_ELEMS; }
  __host__ __device__ int64_t elems() const { return nt * (int64_t)tpr * SPK_TILE_ELEMS; }
  // element (i,j), |i-j| <= 8*kt+7 assumed in band storage range
  __host__ __device__ int64_t elem_off(int64_t i, int64_t j) const {
    return tile_off(i >> 3, j >> 3) + (i & 7) * 8 + (j & 7);
  }
};

struct CsrDev {
  int n = 0; int64_t nnz = 0;
  int *ia = nullptr, *ja = nullptr; double *a = nullptr;
};

struct spk_ctx {
  spk_opts opts;
  cudaStream_t stream;
  // side stream (capi.cu): the spike tips and reduced blocks of a narrow-band factorisation run here, next to the
  // first solve's partition sweeps on `stream`; the reduced solve joins on ev_join.  nullptr = everything on `stream`.
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int side_pending;     // work recorded in ev_join has not been joined into `stream` yet
  int sm_count;
  BandLayout L;
  int have_band, factored, keep_orig;
  double *band;        // device band, factored in place
  double *orig;        // optional copy of the unfactored band (for spk_mult after factor)
  double *dinv;        // nt tiles: inverse diagonal blocks (Linv strictly-lower | Uinv upper)
  // partitions
  int P;               // partitions
  int tipT;            // truncation window in tile rows (>= 2*kt), <= min partition length
  int64_t *h_pstart;   // host copy, P+1 tile-row boundaries
  int64_t *d_pstart;
  // tips / reduced system: interface i couples partition i (bottom) and i+1 (top), i in [0,P-1)
  // plus one extra slot (index P-1) for the boundary with the right-neighbour rank.
  int kp;              // 8*kt
  double *Sb;          // P   * kp*kp : bottom Schur blocks S_b (row-major kp x kp)
  double *St;          // P   * kp*kp : top Schur blocks from the UL window
  double *Vb;          // P   * kp*kp : V^(b)
  double *Wt;          // P   * kp*kp : W^(t)  (Wt[i] belongs to partition i; partition 0 unused unless rank>0)
  double *Red;         // P   * kp*kp : LU (partial pivoting) of I - Wt[i+1] Vb[i]
  int    *RedPiv;      // P   * kp
  // per-solve work
  double *work;        // n_padded * max_nrhs
  double *gtip, *xtip; // (P+1) * 2 * kp
  double *xb, *xt;     // P * kp each
  double *corr;        // P * 2 * tipT*8
  int64_t work_elems;
  // remote (multi-GPU) boundary buffers
  double *remoteWt, *remoteGtop, *remoteXbot, *xtopRemote, *xbBoundary;   // (the vector items hold kp x bnd_cols doubles)
  double *gtopOut;     // g^(t) of every column packed for the left neighbour (bnd_cols > 1)
  int cur_nrhs;        // columns of the split-phase solve in progress
  double *haloL, *haloR; // MatMult halos (8*kt entries of the neighbours' x)
  double *cur_x;       // output vector of the solve in progress (split-phase)
  int have_remote_wt;  // the right neighbour's W^(t) has been set for the factorisation in progress
  int boundary_done;   // the boundary reduced block has been factored
  int wt_done;         // every W^(t) of the factorisation in progress has been computed (factor phase 10)
  // NVLink peer mailboxes (peer.cu): mine, the neighbours' (0 left, 1 right), per-channel sequence numbers
  double *mbox, *peer_mbox[2];
  int peer_ipc[2];
  unsigned long long peer_seq_out[3], peer_seq_in[3];
  unsigned long long* h_peer_err;   // pinned host mirror of the mailbox error word (peer.cu)
  double *tips_mr, *work_mr; int nrhs_mr;   // multi-right-hand-side scratch (grow-only): coupling right-hand sides, sweep results
  // wide-band path (wide.cuh): half-bandwidth > 128, super-blocks of 8x8 tiles
  int wide;                  // 1: the band uses the wide factor format and kernels
  int kb;                    // band width in super-blocks per side (even, 4..8)
  int wide_G;                // column CTAs per partition group of the wide LU (0 = kb)
  unsigned long long* wide_flags; int wide_flag_parts;   // dataflow flags of the wide LU (WIDE_FLAGS_PER_PART words per partition)
  unsigned int* wide_abort;  // set by a wide kernel whose bounded wait expired
  double* wide_zero;         // 4 KB of zeros: the factor run of a window row outside a sweep job (wide_sweep.cu)
  double* wband;             // row/column-reversed copies of the top tip windows (P partitions of tipT tile rows), factored for W^(t)
  double* rband;             // reduced matrices I - W V in band format (P partitions of 8*kb tile rows), factored for R
  double* VbT;               // V^(b) transposed (right operand of the reduced-matrix product)
  int64_t *d_wpstart, *d_rpstart;   // partition boundaries (tile rows) inside wband / rband
  void* d_wjobs; int wjobs_cap;     // device arrays of sweep jobs, one per call site
  void* h_wjobs;                    // host copies of what those arrays hold (upload only on change)
  double* redw; int redw_cols;      // scratch of the wide reduced solve: g_b, g_t, t, x_t, x_b (kp x columns per interface)
  int bnd_cols;                     // right-hand-side columns the boundary exchange buffers (remoteGtop, remoteXbot, xbBoundary) hold
  double *rscale, *cscale;   // optional equilibration (spk_set_scaling): the factored band is diag(r) A diag(c)
  double *cscale_base;       // allocation behind cscale: [kp left-halo scales | n local | kp right-halo scales]
  void* stage[4]; size_t stage_bytes[4];   // grow-only staging for host-vector calls (capi.cu)
  double* kry_ws; size_t kry_ws_bytes;     // grow-only Krylov workspace (krylov.cu: dot partials + basis / work vectors), released with the band
  // operator for Krylov
  CsrDev opA;
  // bookkeeping
  double anorm_max, frac;
  int64_t *d_boost;    // device counter
  double *d_scalar;    // small device scratch (64 doubles)
  int64_t boosted;
  float factor_ms, solve_ms;
  int launches;
  cudaEvent_t ev0, ev1;     // factor start/stop
  cudaEvent_t evs0, evs1;   // solve start/stop
  int timed_factor, timed_solve;
  int timing;               // spk_set_timing: record the factor / solve / per-stage events (each costs ~2 us between two kernels)
  cudaEvent_t evst[8][2];   // per-stage start/stop (see spk_info.stage_ms)
  int stage_timed[8];
  void* lu_trace;           // debug: device buffer for clock64 stamps of the LU kernel (tools only)
  char err[512];
};

void spk_peer_release(spk_ctx* c);   // peer.cu
int spk_peer_failed(spk_ctx* c);     // an earlier exchange expired (pinned mirror, no sync)
int spk_peer_note(spk_ctx* c);       // enqueue the mirror copy of the error word
extern "C" void spk_side_join(spk_ctx* c);   // capi.cu: make c->stream wait for the side-stream work of the last factorisation
// band the LU and the tip windows read: the kept original if there is one that still equals the unfactored band
static inline double* spk_lu_source(spk_ctx* c) { return (c->orig && !c->rscale) ? c->orig : c->band; }
int spk_bnd_desc(spk_ctx* c, int which, double** ptr, size_t* count, int* is_out);   // capi.cu

#define SPK_SET_ERR(ctx, ...) do { if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); } while (0)
#define SPK_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    SPK_SET_ERR(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); return SPK_ERR_CUDA; } } while (0)
#define SPK_KERNEL_CHECK(ctx) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { \
    SPK_SET_ERR(ctx, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); return SPK_ERR_CUDA; } \
    (ctx)->launches++; } while (0)

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spk_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double spk_u01(uint64_t seed, uint64_t counter) {
  return (double)(spk_splitmix64(seed ^ counter) >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// D(8x8) += A(8x4) * B(4x8), fp64 tensor core (SASS DMMA.8x8x4).
// Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + 0/1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t* bar, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch ------------------------------------------------------------
// pdl_trigger(): the next kernel of the stream may start launching once every CTA of this grid has got here;
// pdl_wait(): block until the previous kernel of the stream has completed and its memory is visible (a no-op for a
// kernel launched without the attribute).  Nothing a predecessor writes may be touched before pdl_wait().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef __CUDACC__
static inline bool spk_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SPIKE_B200_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute
template <class... KArgs, class... Args>
static inline cudaError_t spk_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = spk_pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// ---- kernels' host launchers (defined in the .cu files) ---------------------------------------
int spk_launch_generate(spk_ctx* c, uint64_t seed, double delta);
int spk_launch_pack_dense(spk_ctx* c, const double* src_dev, int layout);
int spk_launch_pack_rows_chunk(spk_ctx* c, const double* src_dev, int64_t row0, int64_t nrows);
int spk_launch_pack_finish(spk_ctx* c);
int spk_launch_pack_csr(spk_ctx* c, const CsrDev& A, const int* rowperm_dev, const int* icolperm_dev);
int spk_launch_unpack_rows(spk_ctx* c, const double* band, double* rows_dev);
int spk_launch_absmax(spk_ctx* c, const double* band, double* out_dev);
int spk_launch_scale_band(spk_ctx* c, const double* rs_dev, const double* cs_dev);
int spk_launch_vec_scale(spk_ctx* c, double* out, const double* in, const double* s, int64_t n);
int spk_launch_matmult(spk_ctx* c, const double* band, const double* x, double* y);
int spk_launch_lu(spk_ctx* c);            // per-partition LU (+ S_b capture, dinv)
int spk_launch_ul_tips(spk_ctx* c);       // UL window -> S_t
int spk_launch_tips(spk_ctx* c, int what, int unused);  // 0: local tips + reduced blocks, 1: boundary reduced block
int spk_launch_rtop_left(spk_ctx* c, double* rtop, size_t tip_stride, int nrhs);   // column r: rtop + r*tip_stride <- C_0 (remoteXbot + r*kp)
int spk_launch_sweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld);   // g = D^-1 b
int spk_launch_msweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld);   // g = D^-1 b, nrhs columns at once (msweep.cu)
int spk_launch_reduced_solve(spk_ctx* c, double* x, int nrhs, int64_t ld, int iface_lo, int iface_hi);
int spk_launch_corrections(spk_ctx* c, double* x, int nrhs, int64_t ld);
int spk_launch_reduced_solve_multi(spk_ctx* c, double* x, int nrhs, int64_t ld, double* tips);
int spk_launch_mcorrections(spk_ctx* c, double* x, int nrhs, int64_t ld, const double* tips, double* work, int64_t ld_work);
int spk_launch_gather(spk_ctx* c, const int* idx_dev, int inverse, const double* in, double* out, int64_t n);
int spk_launch_csr_mult(spk_ctx* c, const CsrDev& A, const double* x, double* y);
// wide-band path (wide.cu): the spk_launch_* entry points above forward to these when c->wide
int spk_wide_ul_windows(spk_ctx* c);
int spk_wide_band_lu(spk_ctx* c);
int spk_wide_tips(spk_ctx* c, int what);
int spk_wide_main_sweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld);
int spk_wide_reduced_solve(spk_ctx* c, const double* x, int nrhs, int64_t ld, double* rtop, double* rbot, size_t tip_stride);
int spk_wide_corrections(spk_ctx* c, double* x, int nrhs, int64_t ld, const double* rtop, const double* rbot, size_t tip_stride,
                         double* work, int64_t ldw);
