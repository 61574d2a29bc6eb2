// What does a "service" warp pay per instruction kind while the other warps of the SM stream DMMAs the
// way k_band_lu's column warps do (13 warps, 13 independent accumulator tiles, 2 dependent DMMAs each)?
// The chain warp runs straight-line code (64 ops unrolled, outer loop amortised) of one kind:
//   0 DFMA chain              1 SHFL + DFMA            2 MUFU.RCP64H + DFMA     3 DFMA + untaken uniform branch
//   4 STS + LDS + DFMA        5 FFMA chain             6 SHFL + FFMA            7 DMMA dependent chain
//   8 gj8 fp64 with branch    9 gj8 fp64 branch-free  10 gj8 fp32              11 DFMA + taken uniform branch
// usage: microbench4 ; prints one JSON line per (mode, placement, ctas/SM, load on/off)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int NW = 14;   // 13 DMMA warps + 1 chain warp

__device__ __forceinline__ void gj8_branch(double (&row)[8], int r8, double thr, int& nboost) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double piv = __shfl_sync(0xffffffffu, row[k], k);
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(piv));
    const bool isp = (r8 == k);
    const double q = isp ? 0.0 : row[k] * r0;
    const double e = fma(-piv, r0, 1.0);
    const double t = fma(e, e, e);
    double f = fma(q, t, q);
    double rc = fma(r0, t, r0);
    if (fabs(piv) < thr) {
      rc = (piv < 0.0) ? -1.0 / thr : 1.0 / thr;
      f = isp ? 0.0 : row[k] * rc;
      ++nboost;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c == k) continue;
      const double u = __shfl_sync(0xffffffffu, row[c], k);
      row[c] = isp ? u * rc : fma(-f, u, row[c]);
    }
    row[k] = isp ? rc : -f;
  }
}
__device__ __forceinline__ void gj8_nobranch(double (&row)[8], int r8, double thr, double rthr, int& nboost) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double piv = __shfl_sync(0xffffffffu, row[k], k);
    const bool boost = fabs(piv) < thr;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(piv));
    const bool isp = (r8 == k);
    const double e = fma(-piv, r0, 1.0);
    const double t = fma(e, e, e);
    double rc = fma(r0, t, r0);
    rc = boost ? (piv < 0.0 ? -rthr : rthr) : rc;
    nboost += boost ? 1 : 0;
    const double f = isp ? 0.0 : row[k] * rc;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c == k) continue;
      const double u = __shfl_sync(0xffffffffu, row[c], k);
      row[c] = isp ? u * rc : fma(-f, u, row[c]);
    }
    row[k] = isp ? rc : -f;
  }
}
__device__ __forceinline__ void gj8_f32(float (&row)[8], int r8) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float piv = __shfl_sync(0xffffffffu, row[k], k);
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(piv));
    const float rc = fmaf(r0, fmaf(-piv, r0, 1.0f), r0);
    const bool isp = (r8 == k);
    const float f = isp ? 0.0f : row[k] * rc;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c == k) continue;
      const float u = __shfl_sync(0xffffffffu, row[c], k);
      row[c] = isp ? u * rc : fmaf(-f, u, row[c]);
    }
    row[k] = isp ? rc : -f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(NW * 32, 2) k(long long* out, double* sink, int dmma_iters, int outer, int cw, double thr) {
  __shared__ double sm[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == cw) {
    double x = 1.0 + lane * 1e-9;
    float xf = 1.0f + lane * 1e-6f;
    double row[8];
    float rowf[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { row[c] = ((lane & 7) == c ? 4.0 : 0.25) + 1e-3 * lane; rowf[c] = (float)row[c]; }
    int nb = 0;
    for (int i = 0; i < 500; ++i) x = fma(x, 1.0000001, 1e-9);   // let the DMMA warps get going
    const long long t0 = clock64();
    for (int o = 0; o < outer; ++o) {
      if (MODE == 8) { gj8_branch(row, lane & 7, thr, nb); gj8_branch(row, lane & 7, thr, nb); }
      else if (MODE == 9) { gj8_nobranch(row, lane & 7, thr, 1.0 / thr, nb); gj8_nobranch(row, lane & 7, thr, 1.0 / thr, nb); }
      else if (MODE == 10) { gj8_f32(rowf, lane & 7); gj8_f32(rowf, lane & 7); }
      else {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          if (MODE == 0) x = fma(x, 1.0000001, 1e-9);
          if (MODE == 1) { x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = fma(x, 1.0000001, 1e-9); }
          if (MODE == 2) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = fma(r, 1.0000001, 1.0); }
          if (MODE == 3) { x = fma(x, 1.0000001, 1e-9); if (x == thr) { x = 1.0 / (x + 3.0); sink[1] = x; } }
          if (MODE == 4) { *(volatile double*)&sm[lane] = x; __syncwarp(); x = *(volatile double*)&sm[(lane + 1) & 31]; x = fma(x, 1.0000001, 1e-9); __syncwarp(); }
          if (MODE == 5) xf = fmaf(xf, 1.0000001f, 1e-9f);
          if (MODE == 6) { xf = __shfl_sync(0xffffffffu, xf, (lane + 1) & 31); xf = fmaf(xf, 1.0000001f, 1e-9f); }
          if (MODE == 7) { double c1 = 0.0; dmma884(x, c1, 1.0000001, 1e-9); x += c1 * 1e-30; }
          if (MODE == 11) { x = fma(x, 1.0000001, 1e-9); if (x != thr) { asm volatile("" ::: "memory"); x = x + 1e-12; } }
        }
      }
    }
    const long long t1 = clock64();
    double s = x + xf;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += row[c] + rowf[c];
    if (s == 1.2345 || nb == 123456) sink[0] = s;
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else {
    double2 acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = make_double2(i, -i);
    const long long t0 = clock64();
    for (int it = 0; it < dmma_iters; ++it) {
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        dmma884(acc[i].x, acc[i].y, 1.0000001, 1e-9);
        dmma884(acc[i].x, acc[i].y, 1.0000002, 1e-9);
      }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i) s += acc[i].x + acc[i].y;
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == (cw == 0 ? 32 : 0) && blockIdx.x == 0) out[1] = t1 - t0;
  }
}

template <int MODE>
void run(long long* d, double* s, const char* name, int ops_per_outer) {
  for (int load = 0; load < 2; ++load)
    for (int cpsm = 1; cpsm <= 2; ++cpsm)
      for (int cw : {NW - 1, 0}) {
        const int outer = (MODE >= 8 && MODE <= 10) ? 40 : 40;
        const int dmma_iters = load ? 4000 : 0;
        k<MODE><<<148 * cpsm, NW * 32>>>(d, s, dmma_iters, outer, cw, 1e-30);
        CK(cudaDeviceSynchronize());
        long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        printf("{\"mode\":%d,\"name\":\"%s\",\"dmma_load\":%d,\"ctas_per_sm\":%d,\"chain_warp\":%d,\"cycles_per_op\":%.1f,\"dmma_cycles_per_instr_per_warp\":%.1f}\n",
               MODE, name, load, cpsm, cw, h[0] / (double)(outer * ops_per_outer), load ? h[1] / (double)(dmma_iters * 26) : 0.0);
        fflush(stdout);
      }
}
int main() {
  long long* d; double* s; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&s, 64));
  run<0>(d, s, "dfma", 64);
  run<1>(d, s, "shfl+dfma", 64);
  run<2>(d, s, "rcp64h+dfma", 64);
  run<3>(d, s, "dfma+untaken_branch", 64);
  run<11>(d, s, "dfma+taken_branch", 64);
  run<4>(d, s, "sts+lds+dfma", 64);
  run<5>(d, s, "ffma", 64);
  run<6>(d, s, "shfl+ffma", 64);
  run<7>(d, s, "dmma_dep+dadd", 64);
  run<8>(d, s, "gj8_fp64_branch", 2);
  run<9>(d, s, "gj8_fp64_nobranch", 2);
  run<10>(d, s, "gj8_fp32", 2);
  return 0;
}
