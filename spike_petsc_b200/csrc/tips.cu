// tips.cu -- spike tips and reduced system (dense kp x kp work, one CTA per partition/interface).
//   V_i^(b)   = S_b(i)^-1 B_i        (S_b = trailing Schur block left by the LU of partition i)
//   W_i^(t)   = S_t(i)^-1 C_i        (S_t = leading Schur block left by the UL window of partition i)
//   Rinv_i    = (I - W_{i+1}^(t) V_i^(b))^-1      (truncated SPIKE reduced block, explicit inverse)
// [EXTERNAL algorithm: SPIKE (Polizzi/Sameh), SaP::GPU; the reference only names it, README.md:4.]
//
// The dense solves are BLOCKED Gauss-Jordan eliminations on FP64 tensor cores with matrix AND right-hand sides
// resident in shared memory as 8x8 tiles (row-major, 512 B each = DMMA accumulator-fragment order, see lu.cu):
// per block step k (kt = kp/8 of them)
//   (1) warp 0 inverts the pivot tile M(k,k) with the pivot-block inverse of the band LU (Newton-Schulz on the
//       tensor cores from a Jacobi / FP32 Gauss-Jordan start, exact FP64 Gauss-Jordan with the boosting rule as
//       fallback; lu_dev.cuh) -- one step ahead: it updates and inverts M(k+1,k+1) while the other warps run (3);
//   (2) the pivot tile row is scaled, T <- D^-1 T (2 DMMAs per tile), and its transposes are parked in a row
//       buffer so that step (3) reads every operand with one 16 B load per lane;
//   (3) every other tile row is eliminated: T(I,J) -= M(I,k) T(k,J) (2 DMMAs per tile).
// No pivoting across blocks -- these matrices are the Schur blocks the no-pivot band LU itself would go on to
// factor, and I - W V of decaying spikes.  Right-hand-side tile columns that are still zero (the coupling
// blocks are block triangular) are skipped.
#include "lu_dev.cuh"

#define TIPS_THREADS 512
#define TIPS_WARPS (TIPS_THREADS / 32)

// shared-memory carve-up: M[kt][kt], X[kt][ktx], PT[kt+ktx] (transposed scaled pivot row), Dv (pivot inverse)
// wide tips (kt > 14) do not fit M and all kt right-hand-side tile columns in 227 KB: two column passes
__host__ __device__ __forceinline__ int tip_pass_tiles(int kt) { return kt <= 14 ? kt : (kt + 1) / 2; }
struct TipSmem { double* M; double* X; double* PT; double* Dv; };
__device__ __forceinline__ TipSmem tip_carve(double* sm, int kt) {
  TipSmem t;
  const int ktx = tip_pass_tiles(kt);
  t.M = sm; t.X = t.M + (size_t)kt * kt * 64; t.PT = t.X + (size_t)kt * ktx * 64; t.Dv = t.PT + (size_t)(kt + ktx) * 64;   // two tiles
  return t;
}
static size_t tips_smem(int kt) { return sizeof(double) * 64 * ((size_t)kt * kt + (size_t)kt * tip_pass_tiles(kt) + kt + tip_pass_tiles(kt) + 2) + 64; }

// Solve M X = R in place (X <- M^-1 R); M: kt x kt tiles, X: kt x ktx tiles, both tile-major in shared memory.
// xfirst[J] (optional, shared memory, ktx ints): first block row in which right-hand-side tile column J is
// nonzero on entry -- column J then stays zero until block step xfirst[J] and is skipped before it.
__device__ __forceinline__ double2 tip_invert8(const double2& d, int g, int tq, double thr, double rthr) {
  int nboost = 0;
  const double2 dt = cfrag_transpose(d, g, tq);
  double2 x, xt;
  jacobi_start8(d, dt, x, xt, g, tq);
  if (!ns_refine8(d, dt, x, xt, g, tq)) {
    const float2 xf = gj8_f32_cfrag(d, g, tq);
    const float2 xft = cfrag_transpose_f(xf, g, tq);
    x = make_double2(f2d_bits(xf.x), f2d_bits(xf.y));
    xt = make_double2(f2d_bits(xft.x), f2d_bits(xft.y));
    if (!ns_refine8(d, dt, x, xt, g, tq)) x = gj8_cfrag(d, g, tq, thr, rthr, nboost);
  }
  return x;
}

__device__ void dense_solve_tiles(const TipSmem& T, int kt, int ktx, const int* xfirst, double thr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int l0 = 16 * tq + g;
  const double rthr = 1.0 / thr;
  // Dv[k & 1] = inverse of the pivot tile of step k.  Warp 0 works one block step ahead: in step k it first updates
  // the NEXT pivot tile M(k+1,k+1) and inverts it while the other warps eliminate the rest of block column k.
  if (warp == 0) {
    const double2 d = *reinterpret_cast<const double2*>(T.M + 2 * lane);
    *reinterpret_cast<double2*>(T.Dv + 2 * lane) = tip_invert8(d, g, tq, thr, rthr);
  }
  __syncthreads();
  for (int k = 0; k < kt; ++k) {
    // ---- (2) scale the pivot tile row: items = M tiles right of the pivot, then the live X tiles
    const int nm = kt - (k + 1);
    {
      const double2 dinv = *reinterpret_cast<const double2*>(T.Dv + (k & 1) * 64 + 2 * lane);
      for (int it = warp; it < nm + ktx; it += TIPS_WARPS) {
        const bool ism = it < nm;
        const int J = ism ? k + 1 + it : it - nm;
        if (!ism && xfirst && xfirst[J] > k) continue;                       // still a zero tile
        double* tile = ism ? T.M + ((size_t)k * kt + J) * 64 : T.X + ((size_t)k * ktx + J) * 64;
        const double2 tt = make_double2(tile[l0], tile[l0 + 8]);             // C fragment of tile^T
        double2 r = make_double2(0.0, 0.0);
        dmma_cc(r, dinv, tt);                                                // D^-1 * tile
        __syncwarp();
        *reinterpret_cast<double2*>(tile + 2 * lane) = r;
        double* pt = T.PT + (size_t)it * 64;                                 // transposed copy for step (3)
        pt[l0] = r.x; pt[l0 + 8] = r.y;
      }
    }
    __syncthreads();
    // ---- (3) eliminate block column k from all other tile rows; one warp per (row I, item) pair
    const bool look = (k + 1 < kt);
    if (look && warp == 0) {
      // next pivot tile first: M(k+1,k+1) -= M(k+1,k) T'(k,k+1)  (item 0 of row k+1), then its inverse
      double* tile = T.M + ((size_t)(k + 1) * kt + k + 1) * 64;
      const double2 a = neg2(*reinterpret_cast<const double2*>(T.M + ((size_t)(k + 1) * kt + k) * 64 + 2 * lane));
      double2 acc = *reinterpret_cast<const double2*>(tile + 2 * lane);
      dmma_cc(acc, a, *reinterpret_cast<const double2*>(T.PT + 2 * lane));
      *reinterpret_cast<double2*>(tile + 2 * lane) = acc;
      *reinterpret_cast<double2*>(T.Dv + ((k + 1) & 1) * 64 + 2 * lane) = tip_invert8(acc, g, tq, thr, rthr);
    } else {
      // flattened (row, item) pairs dealt round-robin to the warps; (ri, it) advance incrementally, two pairs per
      // trip so that two independent accumulations are in flight
      const int nitem = nm + ktx;
      const int w0 = look ? warp - 1 : warp, nw = look ? TIPS_WARPS - 1 : TIPS_WARPS;
      int ri = 0, it = w0;
      while (it >= nitem) { it -= nitem; ++ri; }
      const int nrow = kt - 1;
      auto item_live = [&](int ri_, int it_) -> bool {
        if (ri_ >= nrow) return false;
        const bool ism = it_ < nm;
        const int J = ism ? k + 1 + it_ : it_ - nm;
        if (!ism && xfirst && xfirst[J] > k) return false;
        const int I = ri_ < k ? ri_ : ri_ + 1;
        if (look && ism && I == k + 1 && J == k + 1) return false;           // warp 0's tile
        return true;
      };
      auto item_tile = [&](int ri_, int it_) -> double* {
        const int I = ri_ < k ? ri_ : ri_ + 1;
        return it_ < nm ? T.M + ((size_t)I * kt + k + 1 + it_) * 64 : T.X + ((size_t)I * ktx + it_ - nm) * 64;
      };
      while (ri < nrow) {
        int ri2 = ri, it2 = it + nw;
        while (it2 >= nitem) { it2 -= nitem; ++ri2; }
        const bool l1 = item_live(ri, it), l2 = item_live(ri2, it2);
        double* t1 = item_tile(ri, it);
        double* t2 = item_tile(ri2 < nrow ? ri2 : ri, ri2 < nrow ? it2 : it);
        const int I1 = ri < k ? ri : ri + 1, I2 = (ri2 < nrow ? (ri2 < k ? ri2 : ri2 + 1) : I1);
        double2 a1, a2, c1, c2, b1, b2;
        if (l1) {
          a1 = neg2(*reinterpret_cast<const double2*>(T.M + ((size_t)I1 * kt + k) * 64 + 2 * lane));
          c1 = *reinterpret_cast<const double2*>(t1 + 2 * lane);
          b1 = *reinterpret_cast<const double2*>(T.PT + (size_t)it * 64 + 2 * lane);
        }
        if (l2) {
          a2 = neg2(*reinterpret_cast<const double2*>(T.M + ((size_t)I2 * kt + k) * 64 + 2 * lane));
          c2 = *reinterpret_cast<const double2*>(t2 + 2 * lane);
          b2 = *reinterpret_cast<const double2*>(T.PT + (size_t)it2 * 64 + 2 * lane);
        }
        if (l1) dmma884(c1.x, c1.y, a1.x, b1.x);
        if (l2) dmma884(c2.x, c2.y, a2.x, b2.x);
        if (l1) { dmma884(c1.x, c1.y, a1.y, b1.y); *reinterpret_cast<double2*>(t1 + 2 * lane) = c1; }   // T(I,J) -= M(I,k) T'(k,J)
        if (l2) { dmma884(c2.x, c2.y, a2.y, b2.y); *reinterpret_cast<double2*>(t2 + 2 * lane) = c2; }
        ri = ri2; it = it2 + nw;
        while (it >= nitem) { it -= nitem; ++ri; }
      }
    }
    __syncthreads();
  }
}

// tile-major shared memory <- row-major kp x kp global matrix (columns c0t.. of it, nct tile columns)
__device__ __forceinline__ void load_tiles(double* dst, int ldt, const double* src, int kp, int kt, int c0t, int nct) {
  for (int e = threadIdx.x; e < kt * nct * 32; e += blockDim.x) {
    const int pr = e & 31, t = e >> 5;            // pair index inside the tile, tile index
    const int I = t / nct, J = t - I * nct;
    const int r = pr >> 2, c = (pr & 3) * 2;
    const double2 v = *reinterpret_cast<const double2*>(src + (size_t)(8 * I + r) * kp + 8 * (c0t + J) + c);
    *reinterpret_cast<double2*>(dst + ((size_t)I * ldt + J) * 64 + 2 * pr) = v;
  }
}
__device__ __forceinline__ void store_tiles(double* dst, int kp, const double* src, int ldt, int kt, int c0t, int nct) {
  for (int e = threadIdx.x; e < kt * nct * 32; e += blockDim.x) {
    const int pr = e & 31, t = e >> 5;
    const int I = t / nct, J = t - I * nct;
    const int r = pr >> 2, c = (pr & 3) * 2;
    *reinterpret_cast<double2*>(dst + (size_t)(8 * I + r) * kp + 8 * (c0t + J) + c) =
        *reinterpret_cast<const double2*>(src + ((size_t)I * ldt + J) * 64 + 2 * pr);
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

// which==0: out[p] = Sb[p]^-1 B_p ; which==1: out[p] = St[p]^-1 C_p
__global__ void __launch_bounds__(TIPS_THREADS) k_spike_tip(const TipArgs a) {
  extern __shared__ __align__(128) double sm[];
  __shared__ int xfirst[SPK_MAX_KT];
  const int KT = a.L.kt, kp = KT * 8;
  const TipSmem T = tip_carve(sm, KT);
  const int p = blockIdx.x + a.first_part;
  const double* S = a.S + (size_t)p * kp * kp;
  double* out = a.out + (size_t)p * kp * kp;
  // right-hand side block straight from the (never overwritten) coupling tiles of the band:
  //   B(I,J) = tile (tb-KT+I, tb+J), inside the band iff J <= I ;  C(I,J) = tile (tb+I, tb-KT+J), iff J >= I
  const int64_t tb = (a.which == 0) ? a.pstart[p + 1] : a.pstart[p];
  const int ktx = tip_pass_tiles(KT);
  for (int c0t = 0; c0t < KT; c0t += ktx) {
    const int nct = (KT - c0t < ktx) ? KT - c0t : ktx;
    load_tiles(T.M, KT, S, kp, KT, 0, KT);
    for (int e = threadIdx.x; e < KT * ktx * 32; e += blockDim.x) {
      const int pr = e & 31, t = e >> 5;
      const int I = t / ktx, Jl = t - I * ktx, J = c0t + Jl;
      double2 v = make_double2(0.0, 0.0);
      if (Jl < nct) {
        if (a.which == 0) { if (J <= I) v = *reinterpret_cast<const double2*>(a.band + a.L.tile_off(tb - KT + I, tb + J) + 2 * pr); }
        else              { if (J >= I) v = *reinterpret_cast<const double2*>(a.band + a.L.tile_off(tb + I, tb - KT + J) + 2 * pr); }
      }
      *reinterpret_cast<double2*>(T.X + ((size_t)I * ktx + Jl) * 64 + 2 * pr) = v;
    }
    // first nonzero block row of each right-hand-side tile column: B: row J ; C: row 0
    if (threadIdx.x < ktx) xfirst[threadIdx.x] = (threadIdx.x < nct) ? (a.which == 0 ? c0t + (int)threadIdx.x : 0) : KT;
    __syncthreads();
    dense_solve_tiles(T, KT, ktx, xfirst, a.thr);
    store_tiles(out, kp, T.X, ktx, KT, c0t, nct);
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
  extern __shared__ __align__(128) double sm[];
  const int kp = a.kp, KT = kp / 8;
  const TipSmem T = tip_carve(sm, KT);
  const int i = blockIdx.x + a.first_iface;
  const double* V = a.Vb + (size_t)i * kp * kp;
  const double* W = (i == a.remote_iface) ? a.remoteWt : a.Wt + (size_t)(i + a.wt_part_offset) * kp * kp;
  double* out = a.Rinv + (size_t)i * kp * kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int ktx = tip_pass_tiles(KT);
  for (int c0t = 0; c0t < KT; c0t += ktx) {
    const int nct = (KT - c0t < ktx) ? KT - c0t : ktx;
    // M = I - W V on the tensor cores
    if (ktx == KT) {
      // stage W in M's space and V in X's space, keep every warp's output tiles in registers until all
      // operands have been read, then overwrite M
      load_tiles(T.M, KT, W, kp, KT, 0, KT);
      load_tiles(T.X, KT, V, kp, KT, 0, KT);
      __syncthreads();
      constexpr int MAXT = (14 * 14 + TIPS_WARPS - 1) / TIPS_WARPS;
      double2 acc[MAXT];
#pragma unroll
      for (int n = 0; n < MAXT; ++n) {
        const int t = warp + TIPS_WARPS * n;
        acc[n] = make_double2(0.0, 0.0);
        if (t < KT * KT) {
          const int I = t / KT, J = t - I * KT;
          acc[n] = make_double2((I == J && g == 2 * tq) ? 1.0 : 0.0, (I == J && g == 2 * tq + 1) ? 1.0 : 0.0);
          for (int q = 0; q < KT; ++q) {
            const double2 wv = *reinterpret_cast<const double2*>(T.M + ((size_t)I * KT + q) * 64 + 2 * lane);
            const double* vt = T.X + ((size_t)q * KT + J) * 64;
            const double2 vv = make_double2(vt[16 * tq + g], vt[16 * tq + g + 8]);   // C fragment of V(q,J)^T
            dmma_cc(acc[n], neg2(wv), vv);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < MAXT; ++n) {
        const int t = warp + TIPS_WARPS * n;
        if (t < KT * KT) *reinterpret_cast<double2*>(T.M + (size_t)t * 64 + 2 * lane) = acc[n];
      }
    } else {
      // wide tips: operands straight from global memory (L2): W tiles as they are, V tiles transposed
      for (int t = warp; t < KT * KT; t += TIPS_WARPS) {
        const int I = t / KT, J = t - I * KT;
        double2 acc = make_double2((I == J && g == 2 * tq) ? 1.0 : 0.0, (I == J && g == 2 * tq + 1) ? 1.0 : 0.0);
        for (int q = 0; q < KT; ++q) {
          const double2 wv = *reinterpret_cast<const double2*>(W + (size_t)(8 * I + g) * kp + 8 * q + 2 * tq);
          const double* vt = V + (size_t)(8 * q) * kp + 8 * J;      // C fragment of V(q,J)^T: rows 2tq, 2tq+1, column g
          const double2 vv = make_double2(vt[(size_t)(2 * tq) * kp + g], vt[(size_t)(2 * tq + 1) * kp + g]);
          dmma_cc(acc, neg2(wv), vv);
        }
        *reinterpret_cast<double2*>(T.M + ((size_t)I * KT + J) * 64 + 2 * lane) = acc;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KT * ktx * 32; e += blockDim.x) {   // right-hand side: identity columns of this pass
      const int pr = e & 31, t = e >> 5;
      const int I = t / ktx, Jl = t - I * ktx;
      const int r = pr >> 2, c = (pr & 3) * 2;
      const bool dg = (Jl < nct) && (I == c0t + Jl);
      *reinterpret_cast<double2*>(T.X + ((size_t)I * ktx + Jl) * 64 + 2 * pr) = make_double2((dg && r == c) ? 1.0 : 0.0, (dg && r == c + 1) ? 1.0 : 0.0);
    }
    __syncthreads();
    dense_solve_tiles(T, KT, ktx, nullptr, a.thr);
    store_tiles(out, kp, T.X, ktx, KT, c0t, nct);
    __syncthreads();
  }
}

// ============================================================================================
// Register-resident variant for kt <= GJ_MAX_KT (the common case; K = 100 -> kt = 13): the same dataflow as the
// band LU (lu.cu).  One warp per tile COLUMN keeps its kt tiles in registers as DMMA accumulators for the whole
// elimination, one more warp inverts pivot tiles one step ahead.  Per block step k the only shared-memory traffic is
// the broadcast of column k (kt tiles, written once by its owner, read by every column warp as the left DMMA
// operand) and of D_k^-1; there are no index loops: the step loop is unrolled, every tile index is a compile-time
// constant.
//   t = D_k^-1 Aug(k,J) (DMMA pair on the register-transposed tile), Aug(I,J) -= Aug(I,k) t for I != k.
// The owner of the next pivot column updates the next pivot row first and hands that tile (the next pivot block) to
// the inverting warp at once, then publishes its finished column for the next step; one CTA-wide barrier per step.
#define GJ_MAX_KT 13
#define GJ_BAR_ALL 1
#define GJ_BAR_D 4        // named barriers 4,5: pivot tile D_k handed to the inverting warp
template <int KT>
struct GjSmem {
  double PK[2][KT][64];   // column k of the augmented matrix as its owner holds it (tile I = Aug(I,k))
  double XC[2][64];       // D_k^-1
  double tD[2][64];       // D_k
};

// the inverting warp (index 2*KT; it never holds a column: called before any accumulator exists)
template <int KT, int NTHR>
__device__ __noinline__ void gj_invert_loop(GjSmem<KT>& S, double thr) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const double rthr = 1.0 / thr;
  for (int k = 0; k < KT; ++k) {
    named_bar_sync(GJ_BAR_D + (k & 1), 64);
    const double2 d = *reinterpret_cast<const double2*>(&S.tD[k & 1][2 * lane]);
    *reinterpret_cast<double2*>(&S.XC[k & 1][2 * lane]) = tip_invert8(d, g, tq, thr, rthr);
    named_bar_sync(GJ_BAR_ALL, NTHR);   // barrier #k
  }
}

// ---- column j of M and column j of X share one warp ------------------------------------------------------
// When the right-hand side block is block triangular (B: X(I,J) = 0 for I < J; C: for I > J; identity) and the
// pivots are taken in the matching order (top-down for B and the identity, bottom-up for C -- the order the UL
// window itself eliminates in), column J of X is untouched until pivot step J, which is exactly the step at which
// column J of M has been published as the pivot column and is never needed again.  The warp that owned M(:,J) then
// loads X(:,J) into the same registers and carries it to the end.  KT column warps + the inverting warp = 448
// threads at 72 registers for kt = 13: two solves per SM, all 295 interfaces of a 296-partition GPU resident at
// once, and every warp has work in every step.
template <int KT, bool REV, class XInit>
__device__ __forceinline__ void gj_eliminate_shared(GjSmem<KT>& S, double2 (&acc)[KT], int col, XInit xinit) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  constexpr int NTHR = (KT + 1) * 32;
  auto give_d = [&](int st, const double2& t) {
    *reinterpret_cast<double2*>(&S.tD[st & 1][2 * lane]) = t;
    named_bar_arrive(GJ_BAR_D + (st & 1), 64);
  };
  constexpr int P0 = REV ? KT - 1 : 0;
  if (col == P0) {
    give_d(0, acc[P0]);
#pragma unroll
    for (int I = 0; I < KT; ++I)
      if (I != P0) *reinterpret_cast<double2*>(&S.PK[0][I][2 * lane]) = acc[I];
  }
#pragma unroll
  for (int st = 0; st < KT; ++st) {
    const int k = REV ? KT - 1 - st : st;          // pivot block of this step (compile-time after unrolling)
    const int kn = REV ? k - 1 : k + 1;            // the next one
    named_bar_sync(GJ_BAR_ALL, NTHR);              // column k of M and D_k^-1 are in shared memory
    if (col == k) xinit(acc, k);                   // my M column is the published pivot column: switch to X(:,k)
    const double2 xc = *reinterpret_cast<const double2*>(&S.XC[st & 1][2 * lane]);
    const double2 ut = cfrag_transpose(acc[k], g, tq);
    double2 w = make_double2(0.0, 0.0);
    dmma_cc(w, ut, xc);                            // t^T = Aug(k,J)^T D^-T
    acc[k] = cfrag_transpose(w, g, tq);
    w = neg2(w);
    const bool next_owner = (st + 1 < KT) && (col == kn);
    const uint32_t pk = smem_u32(&S.PK[st & 1][0][2 * lane]);
    double2 avn = lds_v2(pk + (REV ? (k - 1 + KT) % KT : (k + 1) % KT) * 512);
#pragma unroll
    for (int ii = 1; ii < KT; ++ii) {
      const int I = REV ? (k - ii + KT) % KT : (k + ii) % KT;      // the next pivot row first
      const double2 av = avn;
      if (ii + 1 < KT) avn = lds_v2(pk + (REV ? (k - ii - 1 + 2 * KT) % KT : (k + ii + 1) % KT) * 512);
      dmma_cc(acc[I], av, w);
      if (ii == 1 && next_owner) give_d(st + 1, acc[I]);
    }
    if (next_owner) {
#pragma unroll
      for (int I = 0; I < KT; ++I)
        if (I != kn) *reinterpret_cast<double2*>(&S.PK[(st + 1) & 1][I][2 * lane]) = acc[I];
    }
  }
}

// WHICH = 0: V^(b) = S_b^-1 B (top-down), WHICH = 1: W^(t) = S_t^-1 C (bottom-up)
template <int KT, int WHICH>
__global__ void __launch_bounds__((KT + 1) * 32, 2) k_spike_tip_gj2(const TipArgs a) {
  extern __shared__ __align__(128) double sm[];
  GjSmem<KT>& S = *reinterpret_cast<GjSmem<KT>*>(sm);
  double* Xs = sm + sizeof(GjSmem<KT>) / sizeof(double);          // X(:,j) staged by warp j: tile (I,j) at (j*KT+I)*64
  const int col = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  constexpr int kp = KT * 8;
  const int p = blockIdx.x + a.first_part;
  if (col == KT) { gj_invert_loop<KT, (KT + 1) * 32>(S, a.thr); return; }
  const double* Sg = a.S + (size_t)p * kp * kp;
  double* out = a.out + (size_t)p * kp * kp;
  const int64_t tb = (WHICH == 0) ? a.pstart[p + 1] : a.pstart[p];
  // my column of the coupling block goes to shared memory now (cp.async), into registers at my pivot step
#pragma unroll
  for (int I = 0; I < KT; ++I) {
    const bool nz = (WHICH == 0) ? (col <= I) : (col >= I);
    if (nz) {
      const double* src = a.band + ((WHICH == 0) ? a.L.tile_off(tb - KT + I, tb + col) : a.L.tile_off(tb + I, tb - KT + col)) + 2 * lane;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(Xs + ((size_t)col * KT + I) * 64 + 2 * lane)), "l"(src) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  double2 acc[KT];
#pragma unroll
  for (int I = 0; I < KT; ++I) acc[I] = *reinterpret_cast<const double2*>(Sg + (size_t)(8 * I + g) * kp + 8 * col + 2 * tq);
  auto xinit = [&](double2 (&x)[KT], int k) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int I = 0; I < KT; ++I) {
      const bool nz = (WHICH == 0) ? (k <= I) : (k >= I);
      x[I] = nz ? *reinterpret_cast<const double2*>(Xs + ((size_t)k * KT + I) * 64 + 2 * lane) : make_double2(0.0, 0.0);
    }
  };
  gj_eliminate_shared<KT, WHICH == 1>(S, acc, col, xinit);
#pragma unroll
  for (int I = 0; I < KT; ++I) *reinterpret_cast<double2*>(out + (size_t)(8 * I + g) * kp + 8 * col + 2 * tq) = acc[I];
}

template <int KT>
__global__ void __launch_bounds__((KT + 1) * 32, 2) k_reduced_factor_gj2(const RedArgs a) {
  extern __shared__ __align__(128) double sm[];
  GjSmem<KT>& S = *reinterpret_cast<GjSmem<KT>*>(sm);
  double* Ws = sm + sizeof(GjSmem<KT>) / sizeof(double);
  const int col = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  constexpr int kp = KT * 8;
  const int i = blockIdx.x + a.first_iface;
  const double* V = a.Vb + (size_t)i * kp * kp;
  const double* W = (i == a.remote_iface) ? a.remoteWt : a.Wt + (size_t)(i + a.wt_part_offset) * kp * kp;
  double* out = a.Rinv + (size_t)i * kp * kp;
  load_tiles(Ws, KT, W, kp, KT, 0, KT);
  named_bar_sync(GJ_BAR_ALL, (KT + 1) * 32);
  if (col == KT) { gj_invert_loop<KT, (KT + 1) * 32>(S, a.thr); return; }
  double2 acc[KT];   // column `col` of I - W V
#pragma unroll
  for (int I = 0; I < KT; ++I) acc[I] = make_double2((I == col && g == 2 * tq) ? 1.0 : 0.0, (I == col && g == 2 * tq + 1) ? 1.0 : 0.0);
#pragma unroll 1
  for (int q = 0; q < KT; ++q) {
    const double2 vq = neg2(cfrag_transpose(*reinterpret_cast<const double2*>(V + (size_t)(8 * q + g) * kp + 8 * col + 2 * tq), g, tq));
#pragma unroll
    for (int I = 0; I < KT; ++I) dmma_cc(acc[I], *reinterpret_cast<const double2*>(Ws + ((size_t)I * KT + q) * 64 + 2 * lane), vq);
  }
  auto xinit = [&](double2 (&x)[KT], int k) {   // column k of the identity
#pragma unroll
    for (int I = 0; I < KT; ++I) x[I] = make_double2((I == k && g == 2 * tq) ? 1.0 : 0.0, (I == k && g == 2 * tq + 1) ? 1.0 : 0.0);
  };
  gj_eliminate_shared<KT, false>(S, acc, col, xinit);
#pragma unroll
  for (int I = 0; I < KT; ++I) *reinterpret_cast<double2*>(out + (size_t)(8 * I + g) * kp + 8 * col + 2 * tq) = acc[I];
}

template <int KT>
static cudaError_t launch_tip_gj(spk_ctx* c, int grid, const TipArgs& t) {
  const size_t smem = sizeof(GjSmem<KT>) + sizeof(double) * 64 * KT * KT;
  cudaError_t e;
  if (t.which == 0) {
    e = cudaFuncSetAttribute(k_spike_tip_gj2<KT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_spike_tip_gj2<KT, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) k_spike_tip_gj2<KT, 0><<<grid, (KT + 1) * 32, smem, c->stream>>>(t);
  } else {
    e = cudaFuncSetAttribute(k_spike_tip_gj2<KT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_spike_tip_gj2<KT, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) k_spike_tip_gj2<KT, 1><<<grid, (KT + 1) * 32, smem, c->stream>>>(t);
  }
  return e;
}
template <int KT>
static cudaError_t launch_red_gj(spk_ctx* c, int grid, const RedArgs& r) {
  const size_t smem = sizeof(GjSmem<KT>) + sizeof(double) * 64 * KT * KT;
  cudaError_t e = cudaFuncSetAttribute(k_reduced_factor_gj2<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_reduced_factor_gj2<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_reduced_factor_gj2<KT><<<grid, (KT + 1) * 32, smem, c->stream>>>(r);
  return cudaSuccess;
}
#define GJ_CASES CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
static bool tip_launch(spk_ctx* c, int grid, const TipArgs& t, size_t smem_generic) {
  switch (c->L.kt) {
#define CASE(K_) case K_: return launch_tip_gj<K_>(c, grid, t) == cudaSuccess;
    GJ_CASES
#undef CASE
    default: k_spike_tip<<<grid, TIPS_THREADS, smem_generic, c->stream>>>(t); return true;
  }
}
static cudaError_t red_launch(spk_ctx* c, int grid, const RedArgs& r, size_t smem_generic) {
  switch (c->L.kt) {
#define CASE(K_) case K_: return launch_red_gj<K_>(c, grid, r);
    GJ_CASES
#undef CASE
    default: k_reduced_factor<<<grid, TIPS_THREADS, smem_generic, c->stream>>>(r); return cudaSuccess;
  }
}

// Spike tips for this rank.  Interface i couples partition i (bottom) with partition i+1 (top);
// interface P-1 is the boundary with the right-neighbour rank (its W^(t) arrives in c->remoteWt).
//   what = 0: every local tip and local reduced block
//   what = 1: only the boundary reduced block (after remoteWt has been set)
//   what = 2: only W^(t) of my first partition (needs the tip windows, not the band LU: sent to the left
//             neighbour while the LU runs)
//   what = 3: like 0, and the boundary reduced block in the same launch (remoteWt already set)
//   what = 4: every W^(t) of this rank (they need the tip windows only; what = 0 / 3 then skip them)
int spk_launch_tips(spk_ctx* c, int what, int unused) {
  (void)unused;
  if (c->wide) return spk_wide_tips(c, what);
  const int kp = c->kp, P = c->P;
  const size_t smem = tips_smem(c->L.kt);
  SPK_CUDA(c, cudaFuncSetAttribute(k_spike_tip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SPK_CUDA(c, cudaFuncSetAttribute(k_reduced_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool has_left = c->opts.rank > 0, has_right = c->opts.rank + 1 < c->opts.nranks;
  RedArgs r;
  r.Vb = c->Vb; r.Wt = c->Wt; r.Rinv = c->Red; r.kp = kp; r.wt_part_offset = 1; r.remoteWt = c->remoteWt;
  r.thr = 1e-13;
  if (what == 1) {
    if (!has_right) return SPK_OK;
    r.first_iface = P - 1; r.remote_iface = P - 1;
    SPK_CUDA(c, red_launch(c, 1, r, smem));
    SPK_KERNEL_CHECK(c);
    return SPK_OK;
  }
  TipArgs t;
  t.band = c->band; t.L = c->L; t.pstart = c->d_pstart; t.thr = c->opts.boost_rel * c->anorm_max;
  if (what == 2) {
    if (!has_left) return SPK_OK;
    t.S = c->St; t.out = c->Wt; t.first_part = 0; t.which = 1;
    tip_launch(c, 1, t, smem);
    SPK_KERNEL_CHECK(c);
    return SPK_OK;
  }
  // V^(b) of partitions 0..P-2 (+ P-1 when a right neighbour exists)
  const int nvb = (P - 1) + (has_right ? 1 : 0);
  if (nvb > 0 && what != 4) {
    t.S = c->Sb; t.out = c->Vb; t.first_part = 0; t.which = 0;
    tip_launch(c, nvb, t, smem);
    SPK_KERNEL_CHECK(c);
  }
  // W^(t) of partitions 1..P-1 (+ 0 when a left neighbour exists)
  const int wfirst = has_left ? 0 : 1;
  if (P - wfirst > 0 && !c->wt_done) {
    t.S = c->St; t.out = c->Wt; t.first_part = wfirst; t.which = 1;
    tip_launch(c, P - wfirst, t, smem);
    SPK_KERNEL_CHECK(c);
  }
  if (what == 4) { c->wt_done = 1; return SPK_OK; }
  const int nred = (P - 1) + ((what == 3 && has_right) ? 1 : 0);   // interface P-1 = boundary with the right rank
  if (nred > 0) {
    r.first_iface = 0; r.remote_iface = (what == 3 && has_right) ? P - 1 : -1;
    SPK_CUDA(c, red_launch(c, nred, r, smem));
    SPK_KERNEL_CHECK(c);
  }
  return SPK_OK;
}
