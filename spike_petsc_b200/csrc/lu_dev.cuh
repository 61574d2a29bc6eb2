// lu_dev.cuh -- device helpers shared by the band-LU kernel (lu.cu) and the spike-tip kernels (tips.cu):
// fragment algebra on 8x8 tiles, pivot-block inverses on the tensor cores, kernel arguments.
#pragma once
#include "common.cuh"
#include <cstdlib>

#define LU_R 4            // ring depth of the step packages
#ifndef LU_PUBG
#define LU_PUBG 6         // package tiles announced per mbarrier
#endif
#ifndef LU_NSM_WIDE
#define LU_NSM_WIDE 0
#endif
#define LU_TRACE_STEPS 64
// optional clock64 trace of CTA 0 (tools/lu_trace.py): [step-100][16]; 0..7 column warp 0, 8..15 inverter warp
// LU_TRV: stamp once a register value has been produced
#define LU_TRV(slot, val) do { if (TRACE && blockIdx.x == 0 && s >= 100 && s < 100 + LU_TRACE_STEPS) { double e_; asm volatile("add.f64 %0, %1, %1;" : "=d"(e_) : "d"(val)); if (lane == 0) a.trace[(s - 100) * 16 + (slot)] = clock64() + (e_ == 1.2345e300 ? 1 : 0); } } while (0)
#define LU_TR(slot) do { if (TRACE && blockIdx.x == 0 && lane == 0 && s >= 100 && s < 100 + LU_TRACE_STEPS) a.trace[(s - 100) * 16 + (slot)] = clock64(); } while (0)

// named barriers 4,5: "the three tiles of update(u) are in shared memory", u even/odd
// (two producer warps arrive, the inverter warp syncs)
#define LU_BAR_TILES 4

struct LuArgs {
  double* band;           // the band the tiles are READ from (the kept original when there is one, else = destination)
  long long dst_off;      // destination band - source band, in doubles: factor tiles are written to band + dst_off
                          // (tiles the elimination never rewrites -- band-edge tiles, coupling blocks -- are equal in both)
  double* schur;          // P * kp*kp : S_b (FWD) or S_t (REV)
  const int64_t* pstart;  // P+1 tile-row boundaries
  long long* boost_count;
  int tpr;
  int tipT;               // REV: window length in tile rows
  int first_part;         // REV: first partition index handled by blockIdx 0
  double boost_thr;
  long long* trace;       // optional debug stamps; nullptr in production
};

// ---- 8x8 tiles in registers (lane = 4g + tq) -----------------------------------------------------
// C fragment of M: lane holds M[g][2tq], M[g][2tq+1]   (= row-major doubles 2*lane, 2*lane+1).
// dmma_cc: acc += M1 * M2 with  m1 = C fragment of M1,  m2t = C fragment of M2^T.
__device__ __forceinline__ void dmma_cc(double2& acc, const double2& m1, const double2& m2t) {
  dmma884(acc.x, acc.y, m1.x, m2t.x);
  dmma884(acc.x, acc.y, m1.y, m2t.y);
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {   // volatile 16 B shared-memory load
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double neg_bits(double x) {   // -x on the integer pipe
  return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
__device__ __forceinline__ double2 neg2(const double2& v) { return make_double2(neg_bits(v.x), neg_bits(v.y)); }
// C fragment of M^T from the C fragment of M: lane wants M[2tq][g], M[2tq+1][g]
__device__ __forceinline__ double2 cfrag_transpose(const double2& c, int g, int tq) {
  const int s0 = 8 * tq + (g >> 1), s1 = s0 + 4;
  const double x0 = __shfl_sync(0xffffffffu, c.x, s0), y0 = __shfl_sync(0xffffffffu, c.y, s0);
  const double x1 = __shfl_sync(0xffffffffu, c.x, s1), y1 = __shfl_sync(0xffffffffu, c.y, s1);
  return make_double2((g & 1) ? y0 : x0, (g & 1) ? y1 : x1);
}
__device__ __forceinline__ float2 cfrag_transpose_f(const float2& c, int g, int tq) {
  const int s0 = 8 * tq + (g >> 1), s1 = s0 + 4;
  const float x0 = __shfl_sync(0xffffffffu, c.x, s0), y0 = __shfl_sync(0xffffffffu, c.y, s0);
  const float x1 = __shfl_sync(0xffffffffu, c.x, s1), y1 = __shfl_sync(0xffffffffu, c.y, s1);
  return make_float2((g & 1) ? y0 : x0, (g & 1) ? y1 : x1);
}
// store a C fragment so that the 64 doubles at `dst` hold M^T row-major
__device__ __forceinline__ void store_transposed(double* dst, const double2& c, int g, int tq) {
  dst[(2 * tq) * 8 + g] = c.x;
  dst[(2 * tq + 1) * 8 + g] = c.y;
}

// In-register Gauss-Jordan inverse of an 8x8 block held as a C fragment, no pivoting, diagonal boosting
// (|pivot| < thr -> +-thr, SpikeGPU style).  Per pivot: 4 fp64 shuffles (pivot, this row's multiplier, the
// two pivot-row entries of this lane's columns), reciprocal (hardware seed + one Halley step, relative
// error e^3), 2 FMAs.  Branch-free.
__device__ __forceinline__ double2 gj8_cfrag(double2 v, int g, int tq, double thr, double rthr, int& nboost) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double mine = (k & 1) ? v.y : v.x;
    const double piv = __shfl_sync(0xffffffffu, mine, 4 * k + (k >> 1));    // D[k][k]
    const double colk = __shfl_sync(0xffffffffu, mine, 4 * g + (k >> 1));   // D[g][k]
    const double rx = __shfl_sync(0xffffffffu, v.x, 4 * k + tq);            // D[k][2tq]
    const double ry = __shfl_sync(0xffffffffu, v.y, 4 * k + tq);            // D[k][2tq+1]
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(piv));
    const double e = fma(-piv, r0, 1.0);
    const double t = fma(e, e, e);
    double rc = fma(r0, t, r0);                                             // 1/piv to fp64 accuracy
    const bool boost = fabs(piv) < thr;                                     // warp-uniform, rare
    rc = boost ? (piv < 0.0 ? -rthr : rthr) : rc;
    nboost += boost ? 1 : 0;
    const bool isp = (g == k);
    const double f = isp ? 0.0 : colk * rc;   // multiplier of the pivot row for this lane's row
    const double ck = isp ? rc : -f;          // column k of the inverse-in-progress
    double nx = isp ? rx * rc : fma(-f, rx, v.x);
    double ny = isp ? ry * rc : fma(-f, ry, v.y);
    if (tq == (k >> 1)) { if (k & 1) ny = ck; else nx = ck; }
    v.x = nx; v.y = ny;
  }
  return v;
}

// fp64 <-> fp32 by bit manipulation on the integer pipe (the F2F conversions would queue behind the column
// warps' DMMAs on the FP64 pipe).  Truncating; out-of-range magnitudes become 0 / inf and simply make the
// fp32 attempt below fail its residual test.
__device__ __forceinline__ float d2f_bits(double d) {
  const unsigned hi = (unsigned)__double2hiint(d), lo = (unsigned)__double2loint(d);
  const int e = (int)((hi >> 20) & 0x7ffu) - 896;
  unsigned b = (hi & 0x80000000u);
  if (e >= 255) b |= 0x7f800000u;
  else if (e > 0) b |= ((unsigned)e << 23) | ((hi & 0xfffffu) << 3) | (lo >> 29);
  return __uint_as_float(b);
}
__device__ __forceinline__ double f2d_bits(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned ex = (b >> 23) & 0xffu;
  const unsigned hi = (b & 0x80000000u) | ((ex + 896u) << 20) | ((b & 0x7fffffu) >> 3);
  return (ex == 0u) ? 0.0 : __hiloint2double((int)(ex == 255u ? (hi | 0x7ff00000u) : hi), (int)(b << 29));
}
// |x| >= 2^e2 (or NaN/inf), decided on the exponent field with integer instructions
__device__ __forceinline__ bool mag_ge_pow2(double x, int e2) {
  return (int)(((unsigned)__double2hiint(x) >> 20) & 0x7ffu) >= 1023 + e2;
}

// Inverse of an 8x8 pivot block held as a C fragment.
// Fast path (well-conditioned blocks, i.e. practically always): Gauss-Jordan without pivoting in FP32 on the
// FP32 pipe -- which the DMMA streams of the column warps do not load -- followed by Newton-Schulz
// iterations X <- X + (I - X D) X on the tensor cores (4 DMMAs each; the residual norm squares per step and
// the iteration stops once the next update is below fp64 round-off), so the result is D^-1 to fp64 accuracy.
// Blocks whose fp32 attempt does not contract (tiny / boostable pivots, huge dynamic range) take the exact
// FP64 Gauss-Jordan with the boosting rule.
__device__ __forceinline__ float2 gj8_f32_cfrag(const double2& d, int g, int tq) {
  float vx = d2f_bits(d.x), vy = d2f_bits(d.y);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float mine = (k & 1) ? vy : vx;
    const float piv = __shfl_sync(0xffffffffu, mine, 4 * k + (k >> 1));
    const float colk = __shfl_sync(0xffffffffu, mine, 4 * g + (k >> 1));
    const float rx = __shfl_sync(0xffffffffu, vx, 4 * k + tq);
    const float ry = __shfl_sync(0xffffffffu, vy, 4 * k + tq);
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(piv));
    const bool isp = (g == k);
    const float f = isp ? 0.0f : colk * rc;
    const float ck = isp ? rc : -f;
    float nx = isp ? rx * rc : fmaf(-f, rx, vx);
    float ny = isp ? ry * rc : fmaf(-f, ry, vy);
    if (tq == (k >> 1)) { if (k & 1) ny = ck; else nx = ck; }
    vx = nx; vy = ny;
  }
  return make_float2(vx, vy);
}
// Newton-Schulz refinement X <- X + (I - X D) X of an approximate inverse, carried for X and X^T at once so
// that every operand is a C fragment already in registers (x, xt: X, X^T; d, dt: D, D^T).  The residual norm
// squares per step; stops once the next update is below fp64 round-off.  false: the iteration does not contract.
__device__ __forceinline__ bool ns_refine8(const double2& d, const double2& dt, double2& x, double2& xt, int g, int tq) {
  const double2 eye = make_double2(g == 2 * tq ? 1.0 : 0.0, g == 2 * tq + 1 ? 1.0 : 0.0);
#pragma unroll 1
  for (int it = 0; it < 7; ++it) {
    double2 r = eye;
    dmma_cc(r, neg2(x), dt);                          // R = I - X D
    // residual entries: all < 2^-28 -> this update is the last one; any >= 2^-4 at the start -> not contracting
    const bool big = mag_ge_pow2(r.x, -28) || mag_ge_pow2(r.y, -28);
    const bool huge = mag_ge_pow2(r.x, -4) || mag_ge_pow2(r.y, -4);
    const unsigned mbig = __ballot_sync(0xffffffffu, big), mhuge = __ballot_sync(0xffffffffu, huge);
    if (mhuge != 0u && (it == 0 || it == 6)) return false;
    const double2 xto = xt;
    dmma_cc(x, r, xto);                               // X   += R X
    if (mbig == 0u) return true;
    dmma_cc(xt, xto, r);                              // X^T += X^T R^T   (operand of the next iteration)
  }
  return false;
}
// Starting guess for diagonally dominant pivot blocks, no elimination at all: the first Newton-Schulz iterate
// from the Jacobi guess diag(D)^-1, which can be written entry-wise,  X = (2I - Dg^-1 D) Dg^-1  (one round of
// shuffles for the diagonal entries, hardware reciprocal seeds).  I - X D = (I - Dg^-1 D)^2.
__device__ __forceinline__ void jacobi_start8(const double2& d, const double2& dt, double2& x, double2& xt, int g, int tq) {
  const int sr = 4 * g + (g >> 1);
  const double drx = __shfl_sync(0xffffffffu, d.x, sr), dry = __shfl_sync(0xffffffffu, d.y, sr);
  const double dc0 = __shfl_sync(0xffffffffu, d.x, 9 * tq);        // D[2tq][2tq]
  const double dc1 = __shfl_sync(0xffffffffu, d.y, 9 * tq + 4);    // D[2tq+1][2tq+1]
  const double dr = (g & 1) ? dry : drx;                           // D[g][g]
  double rr, rc0, rc1;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(dr));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc0) : "d"(dc0));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc1) : "d"(dc1));
  const double e0 = (g == 2 * tq) ? 2.0 : 0.0, e1 = (g == 2 * tq + 1) ? 2.0 : 0.0;
  x = make_double2(fma(-d.x, rr, e0) * rc0, fma(-d.y, rr, e1) * rc1);
  xt = make_double2(fma(-dt.x, rc0, e0) * rr, fma(-dt.y, rc1, e1) * rr);
}

