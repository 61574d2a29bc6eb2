// B200 (sm_100a) FP64 micro-benchmarks used to size the SPIKE kernels.
// Measures: DFMA / DMMA issue rates (alone and mixed), dependent-chain latencies
// (DFMA, reciprocal, shuffle, smem+barrier), and HBM streaming bandwidth
// (LDG.128 and cp.async.bulk).  Results go to stdout as JSON lines; the summary
// is copied into profiles/.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- DFMA rate
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- DMMA rate
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double av, double bv) {
  double c[ILP][4];
  double a[4] = {av, av + 1, av + 2, av + 3};
  double b[2] = {bv, bv + 1};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double av, double bv) {
  double c[ILP][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

// Mixed: even warps DFMA, odd warps DMMA (m8n8k4); reports both FMA counts.
__global__ void k_mixed(double* out, int iters, double a, double b) {
  const int warp = threadIdx.x >> 5;
  if (warp & 1) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
  } else {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
  }
}

// ---------------------------------------------------------------- latency chains (1 warp, clock64)
__global__ void k_lat(long long* out, double x0, int n) {
  double x = x0;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64();
  out[0] = t1 - t0;
  // reciprocal chain (full-precision 1/x)
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / (x + 1.5);
  t1 = clock64();
  out[1] = t1 - t0;
  // __drcp_rn chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = __drcp_rn(x + 1.5);
  t1 = clock64();
  out[2] = t1 - t0;
  // shuffle chain (double = 2 shuffles)
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  t1 = clock64();
  out[3] = t1 - t0;
  // DADD chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = x + 1e-9;
  t1 = clock64();
  out[4] = t1 - t0;
  // approx reciprocal: MUFU.RCP64H + 2 Newton steps
  t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double d = x + 1.5;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    x = r;
  }
  t1 = clock64();
  out[5] = t1 - t0;
  if (x == 12345.678) out[7] = 1;
}

// smem publish + __syncthreads round trip with nthreads threads
__global__ void k_sync(long long* out, int n) {
  __shared__ double buf[1024];
  double x = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    buf[threadIdx.x] = x;
    __syncthreads();
    x = buf[(threadIdx.x + 33) % blockDim.x] + 1.0;
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (x == 12345.678) out[1] = 1;
}

// ---------------------------------------------------------------- HBM streaming
__global__ void k_read(const double2* __restrict__ p, size_t n2, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += stride) { double2 a = p[i]; s += a.x + a.y; }
  if (s == 12345.678) out[0] = s;
}
__global__ void k_copy(const double2* __restrict__ p, double2* __restrict__ q, size_t n2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride) q[i] = p[i];
}

// cp.async.bulk streaming read: each CTA owns a contiguous slab and pulls CH-byte chunks
// through a STAGES-deep smem ring; one thread issues, all threads consume (sum).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
template <int CH, int STAGES>
__global__ void k_bulk_read(const char* __restrict__ p, size_t bytes_per_cta, double* out) {
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t full[STAGES];
  const char* base = p + (size_t)blockIdx.x * bytes_per_cta;
  const int nch = (int)(bytes_per_cta / CH);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES && s < nch; ++s) {
      mbar_expect_tx(&full[s], CH);
      bulk_g2s(sm + s * CH, base + (size_t)s * CH, CH, &full[s]);
    }
  }
  double acc = 0;
  for (int c = 0; c < nch; ++c) {
    const int s = c % STAGES;
    mbar_wait(&full[s], (c / STAGES) & 1);
    const double2* v = reinterpret_cast<const double2*>(sm + s * CH);
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x) { double2 a = v[i]; acc += a.x + a.y; }
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < nch) {
      mbar_expect_tx(&full[s], CH);
      bulk_g2s(sm + s * CH, base + (size_t)(c + STAGES) * CH, CH, &full[s]);
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"smem_optin\":%zu}\n", prop.name, prop.multiProcessorCount, clk_khz,
         (size_t)prop.sharedMemPerBlockOptin);
  const int SMS = prop.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  // ---- compute rates
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 32}) {
    for (int rep = 0; rep < 2; ++rep) {
      k_dfma<16><<<SMS, warps * 32>>>(dout, iters, 1.0000001, 1e-9);
    }
    CK(cudaEventRecord(e0));
    k_dfma<16><<<SMS, warps * 32>>>(dout, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double fma = (double)SMS * warps * 32 * 16.0 * iters;
    printf("{\"bench\":\"dfma\",\"warps_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", warps, ms, 2 * fma / ms / 1e9);
  }
  for (int warps : {4, 8, 16}) {
    k_dmma884<8><<<SMS, warps * 32>>>(dout, iters / 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0));
    k_dmma884<8><<<SMS, warps * 32>>>(dout, iters / 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double fma = (double)SMS * warps * 8.0 * (iters / 4) * 256.0;
    printf("{\"bench\":\"dmma_m8n8k4\",\"warps_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", warps, ms, 2 * fma / ms / 1e9);
  }
  for (int warps : {4, 8, 16}) {
    k_dmma1688<8><<<SMS, warps * 32>>>(dout, iters / 8, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0));
    k_dmma1688<8><<<SMS, warps * 32>>>(dout, iters / 8, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double fma = (double)SMS * warps * 8.0 * (iters / 8) * 1024.0;
    printf("{\"bench\":\"dmma_m16n8k8\",\"warps_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", warps, ms, 2 * fma / ms / 1e9);
  }
  for (int warps : {4, 8, 16}) {
    k_dmma16816<8><<<SMS, warps * 32>>>(dout, iters / 16, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0));
    k_dmma16816<8><<<SMS, warps * 32>>>(dout, iters / 16, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double fma = (double)SMS * warps * 8.0 * (iters / 16) * 2048.0;
    printf("{\"bench\":\"dmma_m16n8k16\",\"warps_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.2f}\n", warps, ms, 2 * fma / ms / 1e9);
  }
  for (int warps : {8, 16}) {
    k_mixed<<<SMS, warps * 32>>>(dout, iters / 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0));
    k_mixed<<<SMS, warps * 32>>>(dout, iters / 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1);
    double fma_v = (double)SMS * (warps / 2) * 32 * 16.0 * (iters / 4);
    double fma_t = (double)SMS * (warps / 2) * 8.0 * (iters / 4) * 256.0;
    printf("{\"bench\":\"mixed_dfma_dmma\",\"warps_per_sm\":%d,\"ms\":%.4f,\"tflops_dfma\":%.2f,\"tflops_dmma\":%.2f,\"tflops_sum\":%.2f}\n",
           warps, ms, 2 * fma_v / ms / 1e9, 2 * fma_t / ms / 1e9, 2 * (fma_v + fma_t) / ms / 1e9);
  }

  // ---- latencies
  long long* dl; CK(cudaMalloc(&dl, 64 * sizeof(long long)));
  CK(cudaMemset(dl, 0, 64 * sizeof(long long)));
  const int n = 4096;
  k_lat<<<1, 32>>>(dl, 0.5, n);
  k_lat<<<1, 32>>>(dl, 0.5, n);
  CK(cudaDeviceSynchronize());
  long long hl[8]; CK(cudaMemcpy(hl, dl, sizeof(hl), cudaMemcpyDeviceToHost));
  printf("{\"bench\":\"latency_cycles\",\"dfma\":%.1f,\"div\":%.1f,\"drcp_rn\":%.1f,\"shfl_f64\":%.1f,\"dadd\":%.1f,\"rcp_approx_2newton\":%.1f}\n",
         hl[0] / (double)n, hl[1] / (double)n, hl[2] / (double)n, hl[3] / (double)n, hl[4] / (double)n, hl[5] / (double)n);
  for (int th : {32, 64, 128, 192, 256, 384, 512}) {
    k_sync<<<1, th>>>(dl, 2048);
    k_sync<<<1, th>>>(dl, 2048);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hl, dl, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("{\"bench\":\"sts_sync_lds_sync_cycles\",\"threads\":%d,\"cycles\":%.1f}\n", th, hl[0] / 2048.0);
  }

  // ---- HBM
  const size_t bytes = (size_t)8 << 30;
  char *p, *q; CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&q, bytes));
  CK(cudaMemset(p, 1, bytes)); CK(cudaMemset(q, 0, bytes));
  for (int bpsm : {4, 8, 16}) {
    k_read<<<SMS * bpsm, 256>>>((const double2*)p, bytes / 16, dout);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; ++r) k_read<<<SMS * bpsm, 256>>>((const double2*)p, bytes / 16, dout);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1) / 3;
    printf("{\"bench\":\"hbm_read_ldg128\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"gbs\":%.1f}\n", bpsm, ms, bytes / ms / 1e6);
  }
  {
    k_copy<<<SMS * 16, 256>>>((const double2*)p, (double2*)q, bytes / 16);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; ++r) k_copy<<<SMS * 16, 256>>>((const double2*)p, (double2*)q, bytes / 16);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms = time_ms(e0, e1) / 3;
    printf("{\"bench\":\"hbm_copy_ldg128\",\"ms\":%.3f,\"gbs_rw\":%.1f}\n", ms, 2.0 * bytes / ms / 1e6);
  }
  {
    constexpr int CH = 8192, ST = 4;
    for (int cps : {1, 2, 4}) {
      const int grid = SMS * cps;
      size_t per = (bytes / grid) / CH * CH;
      CK(cudaFuncSetAttribute(k_bulk_read<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
      k_bulk_read<CH, ST><<<grid, 128, CH * ST>>>(p, per, dout);
      CK(cudaEventRecord(e0));
      for (int r = 0; r < 3; ++r) k_bulk_read<CH, ST><<<grid, 128, CH * ST>>>(p, per, dout);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms = time_ms(e0, e1) / 3;
      printf("{\"bench\":\"hbm_read_bulk8k_x4\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"gbs\":%.1f}\n", cps, ms, (double)per * grid / ms / 1e6);
    }
  }
  {
    constexpr int CH = 16384, ST = 6;
    for (int cps : {1, 2}) {
      const int grid = SMS * cps;
      size_t per = (bytes / grid) / CH * CH;
      CK(cudaFuncSetAttribute(k_bulk_read<CH, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
      k_bulk_read<CH, ST><<<grid, 128, CH * ST>>>(p, per, dout);
      CK(cudaEventRecord(e0));
      for (int r = 0; r < 3; ++r) k_bulk_read<CH, ST><<<grid, 128, CH * ST>>>(p, per, dout);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms = time_ms(e0, e1) / 3;
      printf("{\"bench\":\"hbm_read_bulk16k_x6\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"gbs\":%.1f}\n", cps, ms, (double)per * grid / ms / 1e6);
    }
  }
  CK(cudaDeviceSynchronize());
  printf("{\"done\":true}\n");
  return 0;
}
