// tips.cu -- spike tips and reduced system (dense kp x kp work, one CTA per partition/interface).
//   V_i^(b)   = S_b(i)^-1 B_i        (S_b = trailing Schur block left by the LU of partition i)
//   W_i^(t)   = S_t(i)^-1 C_i        (S_t = leading Schur block left by the UL window of partition i)
//   Rinv_i    = (I - W_{i+1}^(t) V_i^(b))^-1      (truncated SPIKE reduced block, explicit inverse)
// [EXTERNAL algorithm: SPIKE (Polizzi/Sameh), SaP::GPU; the reference only names it, README.md:4.]
// The dense solves are blocked Gauss-Jordan eliminations with matrix AND right-hand sides resident in
// shared memory (2 * kp*(kp+1) doubles: 175 KB at kp = 104, within the 227 KB a B200 CTA may opt into).
#include "common.cuh"

#define TIPS_THREADS 256

// Solve M X = R (kp x kp matrix, ncols right-hand sides), everything resident in shared memory, by
// BLOCKED Gauss-Jordan: kp/8 steps, each inverting an 8x8 pivot block (one warp, same in-register
// Gauss-Jordan + boosting rule as the band LU), scaling the pivot block row and eliminating the block
// column from all other rows with a register-blocked rank-8 update.  No pivoting across blocks --
// these matrices are the Schur blocks the no-pivot band LU itself would go on to factor, and
// I - W V of decaying spikes.  256 threads as a 16 x 16 grid over (rows, columns).
__device__ __forceinline__ void gj8_inverse_warp(const double* D, int ld, double* Dv, double thr) {
  // lanes 0..7 hold rows; result Dv[8][8] row-major
  const int lane = threadIdx.x & 31, r8 = lane & 7;
  double row[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) row[c] = D[r8 * ld + c];
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
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c == k) continue;
      const double u = __shfl_sync(0xffffffffu, row[c], k);
      row[c] = isp ? u * rc : fma(-f, u, row[c]);
    }
    row[k] = isp ? rc : -f;
  }
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) Dv[lane * 8 + c] = row[c];
  }
}

__device__ void dense_solve_smem(double* M, double* X, int ld, int kp, int ncols, int ldx, double* Dv, double thr) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  for (int c0 = 0; c0 < kp; c0 += 8) {
    // (1) inverse of the pivot block
    if (warp == 0) gj8_inverse_warp(M + c0 * ld + c0, ld, Dv, thr);
    __syncthreads();
    // (2) pivot block row <- Dinv * (pivot block row), one thread per column (M columns right of the block, all X columns)
    const int nm = kp - (c0 + 8);
    for (int cc = tid; cc < nm + ncols; cc += TIPS_THREADS) {
      double* col = (cc < nm) ? (M + c0 * ld + (c0 + 8 + cc)) : (X + c0 * ldx + (cc - nm));
      const int st = (cc < nm) ? ld : ldx;
      double v[8], o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = col[k * st];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fma(Dv[r * 8 + k], v[k], acc);
        o[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) col[r * st] = o[r];
    }
    __syncthreads();
    // (3) eliminate the block column from every other row: rank-8 update, thread = rows {ty+16m} x cols {tx+16q}
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      double* T = pass == 0 ? M : X;
      const int tld = pass == 0 ? ld : ldx;
      const int cbeg = pass == 0 ? c0 + 8 : 0;
      const int cend = pass == 0 ? kp : ncols;
      if (cbeg + tx >= cend) continue;
      for (int rbase = ty; rbase < kp; rbase += 16 * 4) {      // 4 rows per register block
        double acc[4][8];
        int rr[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          rr[m] = rbase + 16 * m;
          const bool live = rr[m] < kp && (rr[m] < c0 || rr[m] >= c0 + 8);
          if (!live) rr[m] = -1;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = cbeg + tx + 16 * q;
            acc[m][q] = (live && c < cend) ? T[rr[m] * tld + c] : 0.0;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          double pk[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int c = cbeg + tx + 16 * q; pk[q] = (c < cend) ? T[(c0 + k) * tld + c] : 0.0; }
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const double l = (rr[m] >= 0) ? M[rr[m] * ld + c0 + k] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[m][q] = fma(-l, pk[q], acc[m][q]);
          }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (rr[m] < 0) continue;
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int c = cbeg + tx + 16 * q; if (c < cend) T[rr[m] * tld + c] = acc[m][q]; }
        }
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
  double thr;           // pivot boosting threshold (same rule as the band LU)
};

struct TipSmem { double* M; double* X; double* Dv; };
// wide tips (kp > 112) do not fit M and all kp right-hand sides in 227 KB: they run two column passes
__host__ __device__ __forceinline__ int tip_pass_cols(int kp) { return kp <= 112 ? kp : kp / 2; }
__device__ __forceinline__ TipSmem tip_carve(double* sm, int kp) {
  TipSmem t;
  const int ld = kp + 1, ldx = tip_pass_cols(kp) + 1;
  t.M = sm; t.X = sm + (size_t)kp * ld; t.Dv = t.X + (size_t)kp * ldx;
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
    dense_solve_smem(T.M, T.X, ld, kp, nc, ldx, T.Dv, a.thr);
    for (int r = ty; r < kp; r += 16)
      for (int cc = tx; cc < nc; cc += 16) out[(size_t)r * kp + c0 + cc] = T.X[r * ldx + cc];
    __syncthreads();
  }
}

struct RedArgs {
  const double* Vb; const double* Wt; double* Rinv;
  int kp; int first_iface; int wt_part_offset;  // interface i uses Vb[i], Wt[i + wt_part_offset]
  const double* remoteWt; int remote_iface;     // interface == remote_iface uses remoteWt instead
  double thr;
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
    // M = I - W V: stage W in M's space and V in X's space (coalesced), accumulate each thread's
    // (<= 8 x 8) output patch in registers, then overwrite.  (Wide tips: V does not fit next to M in one
    // pass; they take the slower global-memory product below.)
    if (nc == kp) {
      for (int r = ty; r < kp; r += 16)
        for (int c = tx; c < kp; c += 16) { T.M[r * ld + c] = W[(size_t)r * kp + c]; T.X[r * ldx + c] = V[(size_t)r * kp + c]; }
      __syncthreads();
      double acc[8][8];
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[m][q] = 0.0;
      for (int k = 0; k < kp; ++k) {
        double wv[8], vv[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) { const int r = ty + 16 * m; wv[m] = (r < kp) ? T.M[r * ld + k] : 0.0; }
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int c = tx + 16 * q; vv[q] = (c < kp) ? T.X[k * ldx + c] : 0.0; }
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[m][q] = fma(wv[m], vv[q], acc[m][q]);
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int r = ty + 16 * m, c = tx + 16 * q;
          if (r < kp && c < kp) { T.M[r * ld + c] = ((r == c) ? 1.0 : 0.0) - acc[m][q]; T.X[r * ldx + c] = (r == c) ? 1.0 : 0.0; }
        }
    } else {
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
    }
    __syncthreads();
    dense_solve_smem(T.M, T.X, ld, kp, nc, ldx, T.Dv, a.thr);
    for (int r = ty; r < kp; r += 16)
      for (int cc = tx; cc < nc; cc += 16) out[(size_t)r * kp + c0 + cc] = T.X[r * ldx + cc];
    __syncthreads();
  }
}

static size_t tips_smem(int kp) { return sizeof(double) * ((size_t)kp * (kp + 1) + (size_t)kp * (tip_pass_cols(kp) + 1) + 64) + 64; }

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
  r.thr = 1e-13;
  if (what == 1) {
    if (!has_right) return SPK_OK;
    r.first_iface = P - 1; r.remote_iface = P - 1;
    k_reduced_factor<<<1, TIPS_THREADS, smem, c->stream>>>(r);
    SPK_KERNEL_CHECK(c);
    return SPK_OK;
  }
  TipArgs t;
  t.band = c->band; t.L = c->L; t.pstart = c->d_pstart; t.thr = c->opts.boost_rel * c->anorm_max;
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
