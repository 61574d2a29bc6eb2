// tips.cu -- spike tips and reduced system (dense kp x kp work, one CTA per partition/interface).
//   V_i^(b)   = S_b(i)^-1 B_i        (S_b = trailing Schur block left by the LU of partition i)
//   W_i^(t)   = S_t(i)^-1 C_i        (S_t = leading Schur block left by the UL window of partition i)
//   Rinv_i    = (I - W_{i+1}^(t) V_i^(b))^-1      (truncated SPIKE reduced block, explicit inverse)
// [EXTERNAL algorithm: SPIKE (Polizzi/Sameh), SaP::GPU; the reference only names it, README.md:4.]
// The dense solves use partial pivoting (matrix resident in shared memory); each thread then
// carries one right-hand-side column through the row swaps and the two triangular sweeps.
#include "common.cuh"

#define TIPS_THREADS 256

// LU with partial pivoting of the kp x kp matrix M (shared memory, leading dimension ld).
__device__ void dense_lu_smem(double* M, int ld, int kp, int* piv, double* red_val, int* red_idx) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = 0; j < kp; ++j) {
    // pivot search over rows j..kp-1 of column j
    double best = -1.0; int bi = j;
    for (int r = j + tid; r < kp; r += blockDim.x) {
      const double v = fabs(M[r * ld + j]);
      if (v > best) { best = v; bi = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = red_val[0]; int p = red_idx[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (red_val[w] > b || (red_val[w] == b && red_idx[w] < p)) { b = red_val[w]; p = red_idx[w]; }
      piv[j] = p;
    }
    __syncthreads();
    const int p = piv[j];
    if (p != j) {
      for (int c = tid; c < kp; c += blockDim.x) { const double t = M[j * ld + c]; M[j * ld + c] = M[p * ld + c]; M[p * ld + c] = t; }
    }
    __syncthreads();
    const double d = M[j * ld + j];
    const double rd = (d != 0.0) ? 1.0 / d : 0.0;
    for (int r = j + 1 + tid; r < kp; r += blockDim.x) M[r * ld + j] *= rd;
    __syncthreads();
    const int m = kp - j - 1;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int r = j + 1 + e / m, c = j + 1 + e % m;
      M[r * ld + c] = fma(-M[r * ld + j], M[j * ld + c], M[r * ld + c]);
    }
    __syncthreads();
  }
}
// X (kp x kp, row-major in global memory, column t owned by thread t) <- M^-1 X
__device__ void dense_solve_cols(const double* M, int ld, int kp, const int* piv, double* X) {
  for (int t = threadIdx.x; t < kp; t += blockDim.x) {
    for (int j = 0; j < kp; ++j) {
      const int p = piv[j];
      if (p != j) { const double v = X[(size_t)j * kp + t]; X[(size_t)j * kp + t] = X[(size_t)p * kp + t]; X[(size_t)p * kp + t] = v; }
    }
    for (int r = 1; r < kp; ++r) {
      double s = X[(size_t)r * kp + t];
      for (int j = 0; j < r; ++j) s = fma(-M[r * ld + j], X[(size_t)j * kp + t], s);
      X[(size_t)r * kp + t] = s;
    }
    for (int r = kp - 1; r >= 0; --r) {
      double s = X[(size_t)r * kp + t];
      for (int j = r + 1; j < kp; ++j) s = fma(-M[r * ld + j], X[(size_t)j * kp + t], s);
      X[(size_t)r * kp + t] = s / M[r * ld + r];
    }
  }
}

struct TipArgs {
  const double* band; BandLayout L;
  const int64_t* pstart;
  const double* S;      // Schur blocks (Sb or St), indexed by partition
  double* out;          // Vb or Wt, indexed by partition
  int first_part;       // partition handled by blockIdx 0
  int which;            // 0: Vb (bottom, B block), 1: Wt (top, C block)
};

// which==0: out[p] = Sb[p]^-1 B_p ; which==1: out[p] = St[p]^-1 C_p
__global__ void __launch_bounds__(TIPS_THREADS) k_spike_tip(const TipArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int kp = a.L.kt * 8, ld = kp + 1, KT = a.L.kt;
  double* M = sm;
  int* piv = reinterpret_cast<int*>(M + (size_t)kp * ld);
  double* red_val = reinterpret_cast<double*>(piv + kp + (kp & 1));
  int* red_idx = reinterpret_cast<int*>(red_val + 8);
  const int p = blockIdx.x + a.first_part;
  const double* S = a.S + (size_t)p * kp * kp;
  double* X = a.out + (size_t)p * kp * kp;
  for (int e = threadIdx.x; e < kp * kp; e += blockDim.x) M[(e / kp) * ld + (e % kp)] = S[e];
  // right-hand side block straight from the (never overwritten) coupling tiles of the band
  const int64_t tb = (a.which == 0) ? a.pstart[p + 1] : a.pstart[p];
  for (int e = threadIdx.x; e < kp * kp; e += blockDim.x) {
    const int r = e / kp, c = e % kp;
    double v = 0.0;
    if (a.which == 0) {  // B(r,c) = A(8(tb-KT)+r, 8tb+c), in band iff c/8 <= r/8
      if ((c >> 3) <= (r >> 3)) v = a.band[a.L.elem_off((tb - KT) * 8 + r, tb * 8 + c)];
    } else {             // C(r,c) = A(8tb+r, 8(tb-KT)+c), in band iff c/8 >= r/8
      if ((c >> 3) >= (r >> 3)) v = a.band[a.L.elem_off(tb * 8 + r, (tb - KT) * 8 + c)];
    }
    X[e] = v;
  }
  __syncthreads();
  dense_lu_smem(M, ld, kp, piv, red_val, red_idx);
  __threadfence_block();
  __syncthreads();
  dense_solve_cols(M, ld, kp, piv, X);
}

struct RedArgs {
  const double* Vb; const double* Wt; double* Rinv;
  int kp; int first_iface; int wt_part_offset;  // interface i uses Vb[i], Wt[i + wt_part_offset]
  const double* remoteWt; int remote_iface;     // interface == remote_iface uses remoteWt instead
};
// Rinv[i] = (I - Wt[i+1] Vb[i])^-1
__global__ void __launch_bounds__(TIPS_THREADS) k_reduced_factor(const RedArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int kp = a.kp, ld = kp + 1;
  double* M = sm;
  int* piv = reinterpret_cast<int*>(M + (size_t)kp * ld);
  double* red_val = reinterpret_cast<double*>(piv + kp + (kp & 1));
  int* red_idx = reinterpret_cast<int*>(red_val + 8);
  const int i = blockIdx.x + a.first_iface;
  const double* V = a.Vb + (size_t)i * kp * kp;
  const double* W = (i == a.remote_iface) ? a.remoteWt : a.Wt + (size_t)(i + a.wt_part_offset) * kp * kp;
  double* X = a.Rinv + (size_t)i * kp * kp;
  for (int e = threadIdx.x; e < kp * kp; e += blockDim.x) {
    const int r = e / kp, c = e % kp;
    double s = (r == c) ? 1.0 : 0.0;
    for (int q = 0; q < kp; ++q) s = fma(-W[(size_t)r * kp + q], V[(size_t)q * kp + c], s);
    M[r * ld + c] = s;
    X[e] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  dense_lu_smem(M, ld, kp, piv, red_val, red_idx);
  __threadfence_block();
  __syncthreads();
  dense_solve_cols(M, ld, kp, piv, X);
}

static size_t tips_smem(int kp) { return sizeof(double) * ((size_t)kp * (kp + 1) + 8) + sizeof(int) * (size_t)(kp + 2 + 8) + 64; }

// Spike tips for this rank.  Interface i couples partition i (bottom) with partition i+1 (top);
// interface P-1 is the boundary with the right-neighbour rank (its W^(t) arrives in c->remoteWt).
//   what = 0: every local tip and local reduced block
//   what = 1: only the boundary reduced block (after remoteWt has been set)
int spk_launch_tips(spk_ctx* c, int what, int unused) {
  (void)unused;
  const int kp = c->kp, P = c->P;
  const size_t smem = tips_smem(kp);
  SPK_CUDA(c, cudaFuncSetAttribute(k_spike_tip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SPK_CUDA(c, cudaFuncSetAttribute(k_reduced_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool has_left = c->opts.rank > 0, has_right = c->opts.rank + 1 < c->opts.nranks;
  RedArgs r;
  r.Vb = c->Vb; r.Wt = c->Wt; r.Rinv = c->Red; r.kp = kp; r.wt_part_offset = 1; r.remoteWt = c->remoteWt;
  if (what == 1) {
    if (!has_right) return SPK_OK;
    r.first_iface = P - 1; r.remote_iface = P - 1;
    k_reduced_factor<<<1, TIPS_THREADS, smem, c->stream>>>(r);
    SPK_KERNEL_CHECK(c);
    return SPK_OK;
  }
  TipArgs t;
  t.band = c->band; t.L = c->L; t.pstart = c->d_pstart;
  // V^(b) of partitions 0..P-2 (+ P-1 when a right neighbour exists)
  const int nvb = (P - 1) + (has_right ? 1 : 0);
  if (nvb > 0) {
    t.S = c->Sb; t.out = c->Vb; t.first_part = 0; t.which = 0;
    k_spike_tip<<<nvb, TIPS_THREADS, smem, c->stream>>>(t);
    SPK_KERNEL_CHECK(c);
  }
  // W^(t) of partitions 1..P-1 (+ 0 when a left neighbour exists)
  const int wfirst = has_left ? 0 : 1;
  if (P - wfirst > 0) {
    t.S = c->St; t.out = c->Wt; t.first_part = wfirst; t.which = 1;
    k_spike_tip<<<P - wfirst, TIPS_THREADS, smem, c->stream>>>(t);
    SPK_KERNEL_CHECK(c);
  }
  if (P - 1 > 0) {
    r.first_iface = 0; r.remote_iface = -1;
    k_reduced_factor<<<P - 1, TIPS_THREADS, smem, c->stream>>>(r);
    SPK_KERNEL_CHECK(c);
  }
  return SPK_OK;
}
