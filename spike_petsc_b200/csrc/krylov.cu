// krylov.cu -- device-resident left-preconditioned GMRES(m) / BiCGStab around the SPIKE apply.
// Replaces the inner KSPSolve of KSPSolve_Reorder (/root/reference/src/kspreorder.c:124) with PETSc's
// defaults restated: x0 = 0, left preconditioning, preconditioned residual norm,
// convergence ||r_k|| <= rtol * ||M^{-1} b||, GMRES restart 30 with classical Gram-Schmidt.
// Vectors never leave the GPU; only the (j+1) Hessenberg entries / scalars per iteration do.
#include <cmath>
#include <vector>
#include "common.cuh"

int spk_solve_dev(spk_ctx* c, const double* b, double* x);

// Inner products <V_v, w>, v < nv <= 8, in one pass over w -- DETERMINISTIC: every block reduces its share in a fixed
// order (registers -> warp shuffles -> shared memory, warps in index order) and writes one partial per vector; k_dot_finish
// adds the partials of all blocks in a fixed order.  (The first version accumulated with atomicAdd: the rounding, and with
// it an iteration count at the convergence threshold, depended on the order the blocks happened to finish in.)
__global__ void k_multi_dot(const double* __restrict__ V, int64_t ld, int nv, const double* __restrict__ w, int64_t n,
                            double* __restrict__ partial) {
  __shared__ double sh[8][8];   // [warp][vector]
  double acc[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) acc[v] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
#pragma unroll
    for (int v = 0; v < 8; ++v) if (v < nv) acc[v] = fma(V[(int64_t)v * ld + i], wi, acc[v]);
  }
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    double s = acc[v];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][v] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) s += sh[wp][threadIdx.x];
    partial[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
  }
}
// out[v] = sum over blocks of partial[b][v], one block of 256 threads, fixed order (strided sums, then a shared-memory tree)
__global__ void k_dot_finish(const double* __restrict__ partial, int nblocks, int nv, double* __restrict__ out) {
  __shared__ double sh[256];
  for (int v = 0; v < nv; ++v) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(size_t)b * 8 + v];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = sh[0];
    __syncthreads();
  }
}
// w -= sum_v h[v] V_v
__global__ void k_multi_axpy(const double* __restrict__ V, int64_t ld, int nv, const double* __restrict__ h, double* __restrict__ w, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = w[i];
    for (int v = 0; v < nv; ++v) s = fma(-h[v], V[(int64_t)v * ld + i], s);
    w[i] = s;
  }
}
// y = a*x + b*y   (b may be 0 -> y = a*x without reading y)
__global__ void k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
// z = x + a*y + b*w   (pointers may alias z)
__global__ void k_lin3(const double* x, double a, const double* y, double b, const double* w, double* z, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = x[i] + a * y[i] + (w ? b * w[i] : 0.0);
}

namespace {
struct Kry {
  spk_ctx* c; int64_t n; int grid;
  double* hdev;  // 64 doubles scratch on device (c->d_scalar)
  double* part;  // grid x 8 block partials of the inner products
  int fail = 0;
  void check() { if (cudaGetLastError() != cudaSuccess) fail = 1; c->launches++; }
  int amul(const double* x, double* y) {
    if (c->opA.ia) return spk_launch_csr_mult(c, c->opA, x, y);
    return spk_launch_matmult(c, c->orig, x, y);
  }
  int pc(const double* x, double* y) { return spk_solve_dev(c, x, y); }
  void dots(const double* V, int64_t ld, int nv, const double* w, double* host) {
    for (int v0 = 0; v0 < nv; v0 += 8) {
      const int m = nv - v0 < 8 ? nv - v0 : 8;
      k_multi_dot<<<grid, 256, 0, c->stream>>>(V + (int64_t)v0 * ld, ld, m, w, n, part);
      k_dot_finish<<<1, 256, 0, c->stream>>>(part, grid, m, hdev + v0);
      check();
    }
    cudaMemcpyAsync(host, hdev, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) fail = 1;
  }
  double dot(const double* x, const double* y) { double h; dots(x, n, 1, y, &h); return h; }
  // the same reductions left in device memory at hdev + off (no copy, no synchronisation)
  void dots_dev(const double* V, int64_t ld, int nv, const double* w, int off) {
    for (int v0 = 0; v0 < nv; v0 += 8) {
      const int m = nv - v0 < 8 ? nv - v0 : 8;
      k_multi_dot<<<grid, 256, 0, c->stream>>>(V + (int64_t)v0 * ld, ld, m, w, n, part);
      k_dot_finish<<<1, 256, 0, c->stream>>>(part, grid, m, hdev + off + v0);
      check();
    }
  }
  void fetch(double* host, int cnt) {   // ONE device-to-host copy + synchronisation
    cudaMemcpyAsync(host, hdev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) fail = 1;
  }
  void axpby(double a, const double* x, double b, double* y) { k_axpby<<<grid, 256, 0, c->stream>>>(a, x, b, y, n); check(); }
  void lin3(const double* x, double a, const double* y, double b, const double* w, double* z) { k_lin3<<<grid, 256, 0, c->stream>>>(x, a, y, b, w, z, n); check(); }
};
}  // namespace

int spk_krylov_run(spk_ctx* c, int method, int restart, double rtol, int maxit, const double* b, double* x, int* its,
                   double* rnorm, int* converged) {
  const int64_t n = c->L.n;
  Kry K{c, n, c->sm_count * 8, c->d_scalar, nullptr};
  const int m = restart > 0 ? restart : 30;
  if (m > 60) { SPK_SET_ERR(c, "restart %d too large (max 60)", m); return SPK_ERR_ARG; }
  // workspace kept in the context (grow-only, released with the band): [dot partials | basis or work vectors].  A PC /
  // KSP glue calls this once per right-hand side; a cudaMalloc + cudaFree pair per call costs more than a C1 solve.
  const size_t part_elems = ((size_t)K.grid * 8 + 63) & ~(size_t)63;
  const size_t vec_elems = (size_t)n * (method == SPK_KSP_GMRES ? (size_t)(m + 2) : 7);
  const size_t ws_bytes = sizeof(double) * (part_elems + vec_elems);
  if (c->kry_ws_bytes < ws_bytes) {
    if (c->kry_ws) { cudaStreamSynchronize(c->stream); cudaFree(c->kry_ws); c->kry_ws = nullptr; c->kry_ws_bytes = 0; }
    if (cudaMalloc(&c->kry_ws, ws_bytes) != cudaSuccess) {
      cudaGetLastError();
      SPK_SET_ERR(c, "Krylov workspace of %zu bytes (%s) does not fit", ws_bytes, method == SPK_KSP_GMRES ? "GMRES basis" : "BiCGStab vectors");
      return SPK_ERR_NOMEM;
    }
    c->kry_ws_bytes = ws_bytes;
  }
  K.part = c->kry_ws;
  double* const vecs = c->kry_ws + part_elems;
  int it = 0, conv = 0;
  double res = 0.0;
  int rc = SPK_OK;
  SPK_CUDA(c, cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, c->stream));
  if (method == SPK_KSP_GMRES) {
    double *V = vecs, *t = V + (size_t)n * (m + 1);
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m), hcol(m + 2);
    if ((rc = K.pc(b, V))) goto gdone;
    {
      const double bnorm = std::sqrt(K.dot(V, V));
      res = bnorm;
      if (bnorm == 0.0) { conv = 1; goto gdone; }
      bool first = true;
      while (it < maxit && !conv) {
        if (!first) {  // r = M^{-1}(b - A x)
          if ((rc = K.amul(x, t))) goto gdone;
          K.axpby(1.0, b, -1.0, t);
          if ((rc = K.pc(t, V))) goto gdone;
          res = std::sqrt(K.dot(V, V));
          if (res <= rtol * bnorm) { conv = 1; break; }
        }
        first = false;
        K.axpby(1.0 / res, V, 0.0, V);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = res;
        int j;
        for (j = 0; j < m && it < maxit; ++j) {
          double* vj1 = V + (size_t)(j + 1) * n;
          if ((rc = K.amul(V + (size_t)j * n, t))) goto gdone;
          if ((rc = K.pc(t, vj1))) goto gdone;
          // classical Gram-Schmidt, one host round per iteration: h = V^T w stays on the device and feeds the
          // update w -= V h directly; ||w||^2 lands behind it; (h, ||w||^2) come back in a single copy
          K.dots_dev(V, n, j + 1, vj1, 0);
          k_multi_axpy<<<K.grid, 256, 0, c->stream>>>(V, n, j + 1, K.hdev, vj1, n);
          K.check();
          K.dots_dev(vj1, n, 1, vj1, j + 1);
          K.fetch(hcol.data(), j + 2);
          const double hn = std::sqrt(hcol[j + 1]);
          for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i];
          H[(size_t)(j + 1) * m + j] = hn;
          if (hn != 0.0) K.axpby(1.0 / hn, vj1, 0.0, vj1);
          for (int i = 0; i < j; ++i) {
            const double a0 = H[(size_t)i * m + j], a1 = H[(size_t)(i + 1) * m + j];
            H[(size_t)i * m + j] = cs[i] * a0 + sn[i] * a1;
            H[(size_t)(i + 1) * m + j] = -sn[i] * a0 + cs[i] * a1;
          }
          const double a0 = H[(size_t)j * m + j], a1 = H[(size_t)(j + 1) * m + j], d = std::hypot(a0, a1);
          cs[j] = a0 / d; sn[j] = a1 / d;
          H[(size_t)j * m + j] = d; H[(size_t)(j + 1) * m + j] = 0.0;
          g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j];
          ++it;
          res = std::fabs(g[j + 1]);
          if (res <= rtol * bnorm) { conv = 1; ++j; break; }
        }
        const int jj = j;
        for (int i = jj - 1; i >= 0; --i) {
          double s = g[i];
          for (int q = i + 1; q < jj; ++q) s -= H[(size_t)i * m + q] * y[q];
          y[i] = s / H[(size_t)i * m + i];
        }
        // x += V y  (reuse multi_axpy with negated coefficients)
        for (int i = 0; i < jj; ++i) hcol[i] = -y[i];
        SPK_CUDA(c, cudaMemcpyAsync(K.hdev, hcol.data(), sizeof(double) * jj, cudaMemcpyHostToDevice, c->stream));
        k_multi_axpy<<<K.grid, 256, 0, c->stream>>>(V, n, jj, K.hdev, x, n);
        K.check();
        SPK_CUDA(c, cudaStreamSynchronize(c->stream));
      }
    }
  gdone:
    cudaStreamSynchronize(c->stream);
  } else {
    double *r = vecs, *rh, *p, *v, *s, *t, *tmp;
    rh = r + n; p = rh + n; v = p + n; s = v + n; t = s + n; tmp = t + n;
    SPK_CUDA(c, cudaMemsetAsync(p, 0, sizeof(double) * (size_t)n * 2, c->stream));
    double rho = 1.0, alpha = 1.0, omega = 1.0;
    if ((rc = K.pc(b, r))) goto bdone;
    {
      const double bnorm = std::sqrt(K.dot(r, r));
      res = bnorm;
      SPK_CUDA(c, cudaMemcpyAsync(rh, r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
      if (bnorm == 0.0) conv = 1;
      double rho_next = K.dot(rh, r);
      while (!conv && it < maxit) {
        const double rho1 = rho_next;
        if (rho1 == 0.0) break;
        const double beta = (rho1 / rho) * (alpha / omega);
        // p = r + beta (p - omega v)
        K.lin3(r, beta, p, -beta * omega, v, p);
        if ((rc = K.amul(p, tmp))) goto bdone;
        if ((rc = K.pc(tmp, v))) goto bdone;
        alpha = rho1 / K.dot(rh, v);
        K.lin3(r, -alpha, v, 0.0, nullptr, s);
        if ((rc = K.amul(s, tmp))) goto bdone;
        if ((rc = K.pc(tmp, t))) goto bdone;
        double ts2[2];   // (t,t) and (t,s) in one host round
        K.dots_dev(t, n, 1, t, 0); K.dots_dev(t, n, 1, s, 1); K.fetch(ts2, 2);
        const double tt = ts2[0];
        omega = (tt == 0.0) ? 0.0 : ts2[1] / tt;
        K.lin3(x, alpha, p, omega, s, x);
        K.lin3(s, -omega, t, 0.0, nullptr, r);
        rho = rho1;
        ++it;
        double rr2[2];   // (r,r) for the convergence test and (rh,r) for the next iteration in one host round
        K.dots_dev(r, n, 1, r, 0); K.dots_dev(rh, n, 1, r, 1); K.fetch(rr2, 2);
        res = std::sqrt(rr2[0]); rho_next = rr2[1];
        if (res <= rtol * bnorm) conv = 1;
        if (omega == 0.0) break;
      }
    }
  bdone:
    cudaStreamSynchronize(c->stream);
  }
  if (K.fail && rc == SPK_OK) { SPK_SET_ERR(c, "CUDA failure inside the Krylov loop: %s", cudaGetErrorString(cudaGetLastError())); rc = SPK_ERR_CUDA; }
  *its = it; *rnorm = res; *converged = conv;
  return rc;
}
