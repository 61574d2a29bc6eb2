// FP64 pipe contention: latency of a dependent DFMA chain in one warp while N other warps of the
// same CTA issue DMMA back to back (1 CTA/SM, like the LU kernel).  Also DMMA rate seen by the others.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: chain = DFMA ; 1: chain = SHFL+DFMA ; 2: chain = MUFU.RCP64H + 4 DFMA
template <int ILP>
__global__ void k(long long* out, double* sink, int iters, int chain_iters, int mode, int chain_warp) {
  const int warp = threadIdx.x >> 5;
  if (warp == chain_warp) {
    double x = 1.0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < chain_iters; ++i) {
      if (mode == 0) { x = fma(x, 1.0000001, 1e-9); }
      else if (mode == 1) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = fma(x, 1.0000001, 1e-9); }
      else { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r); x = r + 1.5; }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
    if (x == 1.2345) sink[0] = x;
  } else {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], 1.0000001, 1e-9);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 1.2345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = t1 - t0;
  }
}
int main() {
  long long* d; double* s; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&s, 64));
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {2, 5, 9, 14, 17}) {   // total warps incl. the chain warp
      const int chain_iters = 2000;
      // size DMMA work so that the DMMA warps outlive the chain
      const int iters = 4000;
      k<8><<<148, nw * 32>>>(d, s, iters, chain_iters, mode, nw - 1);
      CK(cudaDeviceSynchronize());
      long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
      const int ops = (mode == 2) ? 5 : 1;
      printf("{\"mode\":%d,\"warps\":%d,\"chain_cycles_per_iter\":%.1f,\"per_fp64_op\":%.1f,\"dmma_cycles_per_instr_per_warp\":%.1f}\n", mode, nw,
             h[0] / (double)chain_iters, h[0] / (double)chain_iters / ops, h[1] / (double)(iters * 8));
    }
  return 0;
}
