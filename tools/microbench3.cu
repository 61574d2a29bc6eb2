// Which warps share an FP64/DMMA pipe?  16-warp CTA, 1 CTA/SM.  `mask` selects warps that stream DMMA,
// `cw` runs a dependent DFMA chain, all other warps exit.  Reports chain cycles/op.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(long long* out, double* sink, int iters, int chain_iters, unsigned mask, int cw) {
  const int warp = threadIdx.x >> 5;
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (warp == cw) {
    double x = 1.0 + threadIdx.x * 1e-9;
    // let the DMMA warps get going
    for (int i = 0; i < 2000; ++i) x = fma(x, 1.0000001, 1e-9);
    long long t0 = clock64();
    for (int i = 0; i < chain_iters; ++i) x = fma(x, 1.0000001, 1e-9);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (x == 1.2345) sink[0] = x;
  } else if (mask & (1u << warp)) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 1.2345) sink[0] = s;
  }
}
int main() {
  long long* d; double* s; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&s, 64));
  struct Case { const char* name; unsigned mask; int cw; };
  Case cases[] = {
    {"none; chain w3", 0x0, 3},
    {"dmma w0; chain w3", 0x1, 3},
    {"dmma w0; chain w4", 0x1, 4},
    {"dmma w0; chain w1", 0x1, 1},
    {"dmma w0; chain w2", 0x1, 2},
    {"dmma w0,1,2; chain w3", 0x7, 3},
    {"dmma w0,4,8; chain w3", 0x111, 3},
    {"dmma w0,4,8; chain w12", 0x111, 12},
    {"dmma w0,4,8; chain w1", 0x111, 1},
    {"dmma w0,4,8,12; chain w15", 0x1111, 15},
    {"dmma all but 3,7,11,15; chain w3", 0x7777, 3},
    {"dmma all but 3,7,11,15; chain w15", 0x7777, 15},
    {"dmma w0..11; chain w15", 0x0FFF, 15},
    {"dmma w0..11; chain w12", 0x0FFF, 12},
    {"dmma w4..15; chain w0", 0xFFF0, 0},
    {"dmma w0..14; chain w15", 0x7FFF, 15},
  };
  for (int ilp : {2, 8})
    for (auto& c : cases) {
      if (ilp == 2) k<2><<<148, 512>>>(d, s, 40000, 4000, c.mask, c.cw);
      else k<8><<<148, 512>>>(d, s, 10000, 4000, c.mask, c.cw);
      CK(cudaDeviceSynchronize());
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      printf("ILP%d  %-36s chain cycles/DFMA = %.1f\n", ilp, c.name, h / 4000.0);
    }
  return 0;
}
