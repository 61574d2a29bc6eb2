// tips.cu -- spike tips and reduced system (dense kp x kp work, one CTA per partition/interface).
//   V_i^(b)   = S_b(i)^-1 B_i        (S_b = trailing Schur block left by the LU of partition i)
//   W_i^(t)   = S_t(i)^-1 C_i        (S_t = leading Schur block left by the UL window of partition i)
//   Rinv_i    = (I - W_{i+1}^(t) V_i^(b))^-1      (truncated SPIKE reduced block, explicit inverse)
// [EXTERNAL algorithm: SPIKE (Polizzi/Sameh), SaP::GPU; the reference only names it, README.md:4.]
// The dense solves use partial pivoting with matrix AND right-hand sides resident in shared memory
// (2 * kp*(kp+1) doubles: 175 KB at kp = 104, within the 227 KB a B200 CTA may opt into).
#include "common.cuh"

#define TIPS_THREADS 256

// Solve M X = R for a kp x kp right-hand side, everything resident in shared memory:
// Gaussian elimination with partial pivoting on the augmented [M | X], then right-looking back
// substitution.  256 threads as a 16 x 16 grid over (rows, columns); no integer division in the loops.
__device__ void dense_solve_smem(double* M, double* X, int ld, int kp, int ncols, int ldx, double* lcol, int* pivs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  for (int j = 0; j < kp; ++j) {
    if (warp == 0) {  // pivot search in column j
      double best = -1.0; int bi = j;
      for (int r = j + lane; r < kp; r += 32) { const double v = fabs(M[r * ld + j]); if (v > best) { best = v; bi = r; } }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) pivs[0] = bi;
    }
    __syncthreads();
    const int p = pivs[0];
    if (p != j) {
      for (int c = j + tid; c < kp; c += TIPS_THREADS) { const double t = M[j * ld + c]; M[j * ld + c] = M[p * ld + c]; M[p * ld + c] = t; }
      for (int c = tid; c < ncols; c += TIPS_THREADS) { const double t = X[j * ldx + c]; X[j * ldx + c] = X[p * ldx + c]; X[p * ldx + c] = t; }
    }
    __syncthreads();
    const double d = M[j * ld + j];
    const double rd = (d != 0.0) ? 1.0 / d : 0.0;
    for (int r = j + 1 + tid; r < kp; r += TIPS_THREADS) lcol[r] = M[r * ld + j] * rd;
    __syncthreads();
    {
      // register-blocked rank-1 update: per row, load the thread's (<= 8) entries, FMA, store back
      double pm[8], px[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int cm = j + 1 + tx + 16 * q, cx = tx + 16 * q;
        pm[q] = (cm < kp) ? M[j * ld + cm] : 0.0;
        px[q] = (cx < ncols) ? X[j * ldx + cx] : 0.0;
      }
      for (int r = j + 1 + ty; r < kp; r += 16) {
        const double l = lcol[r];
        double tm[8], tx8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int cm = j + 1 + tx + 16 * q, cx = tx + 16 * q;
          tm[q] = (cm < kp) ? M[r * ld + cm] : 0.0;
          tx8[q] = (cx < ncols) ? X[r * ldx + cx] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int cm = j + 1 + tx + 16 * q, cx = tx + 16 * q;
          if (cm < kp) M[r * ld + cm] = fma(-l, pm[q], tm[q]);
          if (cx < ncols) X[r * ldx + cx] = fma(-l, px[q], tx8[q]);
        }
      }
    }
    __syncthreads();
  }
  for (int j = kp - 1; j >= 0; --j) {
    const double d = M[j * ld + j];
    const double rd = (d != 0.0) ? 1.0 / d : 0.0;
    for (int c = tid; c < ncols; c += TIPS_THREADS) X[j * ldx + c] *= rd;
    for (int r = tid; r < j; r += TIPS_THREADS) lcol[r] = M[r * ld + j];
    __syncthreads();
    {
      double px[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const int cx = tx + 16 * q; px[q] = (cx < ncols) ? X[j * ldx + cx] : 0.0; }
      for (int r = ty; r < j; r += 16) {
        const double u = lcol[r];
        double t8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int cx = tx + 16 * q; t8[q] = (cx < ncols) ? X[r * ldx + cx] : 0.0; }
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int cx = tx + 16 * q; if (cx < ncols) X[r * ldx + cx] = fma(-u, px[q], t8[q]); }
      }
    }
    __syncthreads();
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

struct TipSmem { double* M; double* X; double* lcol; int* pivs; };
// wide tips (kp > 112) do not fit M and all kp right-hand sides in 227 KB: they run two column passes
__host__ __device__ __forceinline__ int tip_pass_cols(int kp) { return kp <= 112 ? kp : kp / 2; }
__device__ __forceinline__ TipSmem tip_carve(double* sm, int kp) {
  TipSmem t;
  const int ld = kp + 1, ldx = tip_pass_cols(kp) + 1;
  t.M = sm; t.X = sm + (size_t)kp * ld; t.lcol = t.X + (size_t)kp * ldx;
  t.pivs = reinterpret_cast<int*>(t.lcol + kp);
  return t;
}

// which==0: out[p] = Sb[p]^-1 B_p ; which==1: out[p] = St[p]^-1 C_p
__global__ void __launch_bounds__(TIPS_THREADS) k_spike_tip(const TipArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int kp = a.L.kt * 8, ld = kp + 1, KT = a.L.kt;
  const TipSmem T = tip_carve(sm, kp);
  const int p = blockIdx.x + a.first_part;
  const double* S = a.S + (size_t)p * kp * kp;
  double* out = a.out + (size_t)p * kp * kp;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  // right-hand side block straight from the (never overwritten) coupling tiles of the band
  const int64_t tb = (a.which == 0) ? a.pstart[p + 1] : a.pstart[p];
  const int nc = tip_pass_cols(kp), ldx = nc + 1;
  for (int c0 = 0; c0 < kp; c0 += nc) {
    for (int r = ty; r < kp; r += 16) {
      for (int c = tx; c < kp; c += 16) T.M[r * ld + c] = S[(size_t)r * kp + c];
      for (int cc = tx; cc < nc; cc += 16) {
        const int c = c0 + cc;
        double v = 0.0;
        if (a.which == 0) {  // B(r,c) = A(8(tb-KT)+r, 8tb+c), in band iff c/8 <= r/8
          if ((c >> 3) <= (r >> 3)) v = a.band[a.L.elem_off((tb - KT) * 8 + r, tb * 8 + c)];
        } else {             // C(r,c) = A(8tb+r, 8(tb-KT)+c), in band iff c/8 >= r/8
          if ((c >> 3) >= (r >> 3)) v = a.band[a.L.elem_off(tb * 8 + r, (tb - KT) * 8 + c)];
        }
        T.X[r * ldx + cc] = v;
      }
    }
    __syncthreads();
    dense_solve_smem(T.M, T.X, ld, kp, nc, ldx, T.lcol, T.pivs);
    for (int r = ty; r < kp; r += 16)
      for (int cc = tx; cc < nc; cc += 16) out[(size_t)r * kp + c0 + cc] = T.X[r * ldx + cc];
    __syncthreads();
  }
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
  const TipSmem T = tip_carve(sm, kp);
  const int i = blockIdx.x + a.first_iface;
  const double* V = a.Vb + (size_t)i * kp * kp;
  const double* W = (i == a.remote_iface) ? a.remoteWt : a.Wt + (size_t)(i + a.wt_part_offset) * kp * kp;
  double* out = a.Rinv + (size_t)i * kp * kp;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int nc = tip_pass_cols(kp), ldx = nc + 1;
  for (int c0 = 0; c0 < kp; c0 += nc) {
    // M = I - W V (W, V stream from L2; kp^3 flops per pass are negligible next to the elimination)
    for (int r = ty; r < kp; r += 16) {
      const double* wr = W + (size_t)r * kp;
      for (int c = tx; c < kp; c += 16) {
        double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
        int q = 0;
        for (; q + 1 < kp; q += 2) { s0 = fma(-wr[q], V[(size_t)q * kp + c], s0); s1 = fma(-wr[q + 1], V[(size_t)(q + 1) * kp + c], s1); }
        for (; q < kp; ++q) s0 = fma(-wr[q], V[(size_t)q * kp + c], s0);
        T.M[r * ld + c] = s0 + s1;
      }
      for (int cc = tx; cc < nc; cc += 16) T.X[r * ldx + cc] = (r == c0 + cc) ? 1.0 : 0.0;
    }
    __syncthreads();
    dense_solve_smem(T.M, T.X, ld, kp, nc, ldx, T.lcol, T.pivs);
    for (int r = ty; r < kp; r += 16)
      for (int cc = tx; cc < nc; cc += 16) out[(size_t)r * kp + c0 + cc] = T.X[r * ldx + cc];
    __syncthreads();
  }
}

static size_t tips_smem(int kp) { return sizeof(double) * ((size_t)kp * (kp + 1) + (size_t)kp * (tip_pass_cols(kp) + 1) + kp) + sizeof(int) * 4 + 64; }

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
