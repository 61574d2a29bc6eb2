// wide.cuh -- shared definitions of the wide-band path (half-bandwidth 129..512, BASELINE config 5).
//
// Wide bands do not fit the register-resident window of lu.cu (K x K doubles = 2 MB at K = 512), so the
// factorisation is blocked one level up: SUPER-BLOCKS of 8x8 tiles (64 x 64 entries).  With KB = ceil(K/64)
// rounded up to an even number, the band is KB super-blocks wide on either side of the diagonal one.
//
// Storage is the same tile-major band as everywhere else (common.cuh) with kt_store = 8*KB + 7 tiles per side, so
// that EVERY super-block (Ib, Jb), |Ib - Jb| <= KB, is fully addressable: tile row 8*Ib+i holds the 8 tiles of
// super-block column Jb contiguously (one 4 KB run).  The extra 7 tiles are needed because the block factor
// Ub(I, I+KB) = D_I^-1 A~(I, I+KB) fills its whole super-block (the scalar U stays inside the band, the product
// with the explicit 64 x 64 inverse does not).
//
// Factor format (the narrow format one level up):
//     Lb(I,J) = A~(I,J)  (J < I),   diagonal slot = D_I^-1 (explicit 64 x 64 inverse),   Ub(I,J) = D_I^-1 A~(I,J)  (J > I)
// Same Schur complements and scalar pivots as the no-pivot scalar LU (the in-place Gauss-Jordan of D_I meets
// exactly those pivots, so the boosting rule is unchanged).
#pragma once
#include "lu_dev.cuh"

#define WIDE_FLAGS_PER_PART 64      // unsigned long long words per partition in the flag array
#define WIDE_SPIN_LIMIT 4000000000ll // clock64 ticks (~2 s) before a dataflow wait gives up

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double2 ldcg_v2(const double* p) {   // L2 load (data produced by another SM)
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
// whole-warp wait until *flag >= target.  false: gave up (another CTA raised the abort word, or ~2 s passed)
__device__ __forceinline__ bool wide_wait_ge(const unsigned long long* flag, unsigned long long target, unsigned int* abort_word) {
  int ok = 1;
  if ((threadIdx.x & 31) == 0) {
    if (ld_acquire_gpu_u64(flag) < target) {
      const long long t0 = clock64();
      unsigned it = 0;
      while (ld_acquire_gpu_u64(flag) < target) {
        __nanosleep(40);
        if ((++it & 255u) == 0u) {
          if (*reinterpret_cast<volatile unsigned int*>(abort_word) != 0u) { ok = 0; break; }
          if (clock64() - t0 > WIDE_SPIN_LIMIT) { atomicExch(abort_word, 1u); ok = 0; break; }
        }
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}
__device__ __forceinline__ bool wide_ready_ge(const unsigned long long* flag, unsigned long long target) {   // non-blocking probe
  int ok = 0;
  if ((threadIdx.x & 31) == 0) ok = ld_acquire_gpu_u64(flag) >= target;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// ---- host launchers (wide_lu.cu / wide_sweep.cu) ---------------------------------------------------------
struct WideSweepJob {
  const double* band;        // factored band the job sweeps (the context band, the reversed-window band or the reduced band)
  long long sb_lo, sb_hi;    // solve window, in super-block rows of that band
  long long sb_fwd;          // first super-block row of the forward sweep (the right-hand side is zero above it)
  const double* in;          // right-hand side: entry (R, c) at in[(R - row0) * in_rs + c * in_cs]; nullptr = identity columns
  long long in_rs, in_cs;
  double* out;               // y after the forward sweep, x after the backward sweep (may alias in)
  long long out_rs, out_cs;
  long long row0;            // band row of in/out row 0
  long long nrow_valid;      // rows R - row0 in [0, nrow_valid) exist in in/out; others read as zero and are not written
  int ncols;                 // right-hand-side columns
  int pad;
};
int spk_wide_lu(spk_ctx* c, double* band, const int64_t* d_pstart, int P);                 // wide_lu.cu
int spk_wide_sweep(spk_ctx* c, const WideSweepJob* d_jobs, int njobs, int max_cols, bool long_jobs);   // wide_sweep.cu (long_jobs: partition sweeps / corrections)
