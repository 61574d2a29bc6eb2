// wide_lu.cu -- banded block LU for wide bands (K = 129..512) on the FP64 tensor cores.
//
// Replaces PCSetUp(inner) = PETSc sparse LU of the AIJ band (/root/reference/src/matbanded.c:178) for the
// half-bandwidths the register-resident kernel of lu.cu cannot hold.  Factor format: see wide.cuh.
//
// One persistent GROUP of G+1 CTAs per SPIKE partition (G <= KB column CTAs + 1 inverter CTA; groups loop over the
// partitions, the whole grid is co-resident: cooperative launch, one CTA per SM).  Right-looking elimination in
// super-block steps s = 0..T-1 over the KB x KB super-block trailing window, which stays in the band itself
// (L2 resident: 2 MB per partition at K = 512):
//   * column CTA (J mod G) owns super-block column J.  Per step and owned column:
//       U phase:  Ub(s,J) = D_s^-1 A~(s,J)             one 64^3 product; D_s^-1 comes from the inverter CTA
//       C phase:  A~(s+i,J) -= A~(s+i,s) Ub(s,J)       i = 1..KB, one 64^3 product each -- the trailing update,
//                                                      the dense FP64 contraction this kernel exists for
//     Warp r of the CTA owns tile row r of every super-block: its 8 accumulator tiles (a 8 x 64 strip) live in
//     registers, the LEFT operands (tiles of D_s^-1 / of the pivot column A~(.,s)) go from L2 straight into
//     registers as DMMA fragments (a tile row of a super-block is one contiguous 4 KB run; every tile is used for
//     8 products by this warp only), the RIGHT operand Ub(s,J) sits in shared memory as transposed tiles and is
//     shared by the 8 warps.  Tile algebra as in lu.cu: row-major tiles are accumulator fragments, the left
//     operand of M1*M2 is the fragment of M1, the right one the fragment of M2^T.
//   * no grid-wide barrier: three families of release/acquire flags per partition in global memory --
//       dflag[r] : row tile r of the next pivot block D_{s+1} is final           (column owner  -> inverter)
//       dinv     : D_s^-1 is in the band's diagonal slot                         (inverter      -> everybody)
//       lflag[r] : row tile r of pivot column s+1 is final down to row s+1+i     (column owner  -> everybody)
//     The owner of column s+1 updates D_{s+1} first, so the inverter works while the rest of update(s) runs
//     (look-ahead by construction).  Every wait is on an event of an earlier step, so co-residency is all that
//     progress needs; waits are bounded (abort word) so a bug cannot hang the GPU.
//   * inverter CTA: in-place block Gauss-Jordan of the 64 x 64 pivot block in shared memory, 8 x 8 pivot tiles
//     inverted by the tensor-core Newton-Schulz / FP32 / exact-FP64-with-boosting ladder of lu_dev.cuh.
#include "wide.cuh"

struct WideLuArgs {
  double* band; int tpr; int kts;   // storage: tiles per tile row, tiles per side
  int KB, G, P;
  const int64_t* pstart;            // P+1 tile-row boundaries (multiples of 8)
  unsigned long long* flags;        // P * WIDE_FLAGS_PER_PART, zeroed before the launch
  unsigned int* abort_word;
  double boost_thr; long long* boost_count;
  long long* trace;                 // optional (tools/wide_trace.py): [role][8] accumulated clock64 ticks of group 0
};

#define WL_THREADS 256

#define WT_ADD(slot, t0_) do { if (TRACE && group == 0 && lane == 0 && warp == 0) { const long long n_ = clock64(); tr[slot] += n_ - (t0_); (t0_) = n_; } } while (0)
template <bool TRACE>
__global__ void __launch_bounds__(WL_THREADS, 1) k_wide_lu(const WideLuArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const Bt = reinterpret_cast<double*>(smem_raw);   // [8][8][64]  (inverter: the pivot block M)
  double* const UtA = Bt + 4096;                            // [8][8][64]  Ub(s,J) as transposed tiles, double buffered
  double* const UtB = UtA + 4096;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int gsz = a.G + 1;
  const int group = blockIdx.x / gsz, role = blockIdx.x % gsz, ngroups = gridDim.x / gsz;
  const int tpr = a.tpr, kts = a.kts, KB = a.KB;
  auto tile = [&](int64_t I, int64_t J) -> double* { return a.band + (I * tpr + (J - I + kts)) * SPK_TILE_ELEMS; };

  long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = TRACE ? clock64() : 0;
  for (int p = group; p < a.P; p += ngroups) {
    unsigned long long* const fl = a.flags + (size_t)p * WIDE_FLAGS_PER_PART;
    // per warp r: {dinv copy, lflag[r]} adjacent (one 16-byte acquire load probes both), dflag[r] behind them
    unsigned long long* const f_dinv = fl + 2 * warp;
    unsigned long long* const f_l = fl + 2 * warp + 1;
    unsigned long long* const f_d = fl + 16 + warp;
    const int64_t T0 = a.pstart[p];
    const int T = (int)((a.pstart[p + 1] - T0) >> 3);

    if (role == a.G) {
      // =============================== inverter CTA =========================================
      double* const M = Bt;                       // tile (i,j) at M + (i*8+j)*64
      double* const RT = UtA;                     // transposed tiles of the scaled pivot row
      double* const XS = UtA + 8 * 64;            // X = (pivot tile)^-1
      double* const XTS = XS + 64;                // X^T
      const double thr = a.boost_thr, rthr = 1.0 / a.boost_thr;
      int nboost = 0;
      for (int s = 0; s < T; ++s) {
        const int64_t Id = T0 + 8 * s;
        WT_ADD(3, tmark);
        if (s > 0 && !wide_wait_ge(f_d, (unsigned long long)s, a.abort_word)) return;
        WT_ADD(0, tmark);
        {
          const double* src = tile(Id + warp, Id) + 2 * lane;
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(M + (warp * 8 + c) * 64 + 2 * lane) = ldcg_v2(src + c * 64);
        }
        __syncthreads();
        WT_ADD(1, tmark);
#pragma unroll 1
        for (int pv = 0; pv < 8; ++pv) {
          if (warp == pv) {
            const double2 d = *reinterpret_cast<const double2*>(M + (pv * 8 + pv) * 64 + 2 * lane);
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
            xt = cfrag_transpose(x, g, tq);
            *reinterpret_cast<double2*>(XS + 2 * lane) = x;
            *reinterpret_cast<double2*>(XTS + 2 * lane) = xt;
          }
          __syncthreads();
          if (warp != pv) {   // row scale: M[pv][warp] <- X M[pv][warp]
            double* t = M + (pv * 8 + warp) * 64 + 2 * lane;
            const double2 m = *reinterpret_cast<const double2*>(t);
            const double2 mt = cfrag_transpose(m, g, tq);
            const double2 x = *reinterpret_cast<const double2*>(XS + 2 * lane);
            double2 acc = make_double2(0.0, 0.0);
            dmma_cc(acc, x, mt);
            *reinterpret_cast<double2*>(t) = acc;
            *reinterpret_cast<double2*>(RT + warp * 64 + 2 * lane) = cfrag_transpose(acc, g, tq);
          } else {
            *reinterpret_cast<double2*>(M + (pv * 8 + pv) * 64 + 2 * lane) = *reinterpret_cast<const double2*>(XS + 2 * lane);
          }
          __syncthreads();
          if (warp != pv) {   // eliminate tile row `warp`
            double* row = M + (warp * 8) * 64 + 2 * lane;
            const double2 nf = neg2(*reinterpret_cast<const double2*>(row + pv * 64));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j == pv) continue;
              double2 cc = *reinterpret_cast<const double2*>(row + j * 64);
              dmma_cc(cc, nf, *reinterpret_cast<const double2*>(RT + j * 64 + 2 * lane));
              *reinterpret_cast<double2*>(row + j * 64) = cc;
            }
            double2 cp = make_double2(0.0, 0.0);
            dmma_cc(cp, nf, *reinterpret_cast<const double2*>(XTS + 2 * lane));
            *reinterpret_cast<double2*>(row + pv * 64) = cp;
          }
          __syncthreads();
        }
        WT_ADD(2, tmark);
        {
          double* dst = tile(Id + warp, Id) + 2 * lane;
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(dst + c * 64) = *reinterpret_cast<const double2*>(M + (warp * 8 + c) * 64 + 2 * lane);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x < 8) st_release_gpu_u64(fl + 2 * threadIdx.x, (unsigned long long)(s + 1));
      }
      if (lane == 0 && nboost) atomicAdd((unsigned long long*)a.boost_count, (unsigned long long)nboost);
      WT_ADD(3, tmark);
      if (TRACE && group == 0 && threadIdx.x == 0 && a.trace) for (int q = 0; q < 8; ++q) a.trace[role * 8 + q] = tr[q];
      continue;
    }

    // =============================== column CTA ===============================================
    // Flag values are cached per warp (they only grow): a probe costs an acquire load plus an L1 invalidation that
    // waits for every load in flight, so it is made only when the cached value does not already answer, and
    // BEFORE the prefetch loads of the next row are issued.
    int ucount = 0;
    unsigned long long known_l = 0, known_dinv = 0;
    bool have_u = false;            // unext holds my tile row of A~(s,J) (the row updated first in the previous step)
    double2 unext[8];
    for (int s = 0; s < T; ++s) {
      const int nwin = (T - 1 - s) < KB ? (T - 1 - s) : KB;
      const int64_t Is = T0 + 8 * s;                       // first tile row / column of the pivot super-block
      bool kept = false;
      for (int j = 1; j <= nwin; ++j) {
        const int J = s + j;
        if (J % a.G != role) continue;
        double* const Ut = (ucount & 1) ? UtB : UtA;
        ++ucount;
        const int64_t Jc = T0 + 8 * (int64_t)J;            // first tile column of super-block column J
        auto need_flag = [&](int i) -> bool { return s > 0 && i <= KB - 1; };
        auto probe = [&]() {   // one acquire load for both flags
          if (lane == 0) {
            asm volatile("ld.acquire.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(known_dinv), "=l"(known_l) : "l"(f_dinv) : "memory");
          }
          known_dinv = __shfl_sync(0xffffffffu, known_dinv, 0);
          known_l = __shfl_sync(0xffffffffu, known_l, 0);
        };
        auto l_ready = [&](int i) -> bool {                // left operand row s+i final?  (non-blocking)
          if (!need_flag(i)) return true;
          const unsigned long long want = (unsigned long long)(s * 8 + i);
          if (known_l >= want) return true;
          probe();
          return known_l >= want;
        };
        // ---------------- U phase: Ub(s,J) = D_s^-1 A~(s,J) ----------------
        WT_ADD(7, tmark);
        double* const urow = tile(Is + warp, Jc) + 2 * lane;   // my tile row of A~(s,J): 8 contiguous tiles
        if (known_dinv < (unsigned long long)(s + 1)) probe();
        const bool dinv_early = known_dinv >= (unsigned long long)(s + 1);
        const bool l1_early = l_ready(1);
        double2 av[8], cc[8], ccn[8], dv[8];
        if (dinv_early) {
          const double* dsrc = tile(Is + warp, Is) + 2 * lane;
#pragma unroll
          for (int k = 0; k < 8; ++k) dv[k] = ldcg_v2(dsrc + k * 64);
        }
        {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const double2 u = (have_u && j < KB) ? unext[c] : *reinterpret_cast<const double2*>(urow + c * 64);
            store_transposed(Bt + (warp * 8 + c) * 64, u, g, tq);
          }
        }
        // first trailing row: its loads fly during the U phase
        {
          const double* csrc = tile(Is + 8 + warp, Jc) + 2 * lane;
#pragma unroll
          for (int c = 0; c < 8; ++c) cc[c] = *reinterpret_cast<const double2*>(csrc + c * 64);
          if (l1_early) {
            const double* lsrc = tile(Is + 8 + warp, Is) + 2 * lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) av[k] = ldcg_v2(lsrc + k * 64);
          }
        }
        __syncthreads();
        WT_ADD(0, tmark);
        if (!dinv_early) {
          if (!wide_wait_ge(f_dinv, (unsigned long long)(s + 1), a.abort_word)) return;
          known_dinv = (unsigned long long)(s + 1);
          const double* dsrc = tile(Is + warp, Is) + 2 * lane;
#pragma unroll
          for (int k = 0; k < 8; ++k) dv[k] = ldcg_v2(dsrc + k * 64);
        }
        WT_ADD(1, tmark);
        {
          double2 acc[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = make_double2(0.0, 0.0);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            double2 bt[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bt[c] = *reinterpret_cast<const double2*>(Bt + (k * 8 + c) * 64 + 2 * lane);
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma884(acc[c].x, acc[c].y, dv[k].x, bt[c].x);
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma884(acc[c].x, acc[c].y, dv[k].y, bt[c].y);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<double2*>(urow + c * 64) = acc[c];                 // the stored factor Ub(s,J)
            store_transposed(Ut + (warp * 8 + c) * 64, acc[c], g, tq);
          }
        }
        __syncthreads();
        WT_ADD(2, tmark);
        // ---------------- C phase: A~(s+i,J) -= A~(s+i,s) Ub(s,J), i = 1..nwin ----------------
        // left operand rows are final once their producer (the owner of column s, during step s-1) says so;
        // row s+KB was never touched by an update (it enters the window now)
        if (!l1_early) {
          if (need_flag(1) && !wide_wait_ge(f_l, (unsigned long long)(s * 8 + 1), a.abort_word)) return;
          const double* lsrc = tile(Is + 8 + warp, Is) + 2 * lane;
#pragma unroll
          for (int k = 0; k < 8; ++k) av[k] = ldcg_v2(lsrc + k * 64);
        }
        WT_ADD(3, tmark);
        int pending_pub = 0;   // row whose publication (next pivot column only) waits for the stores to drain
        // One trailing row: CC (this row's accumulator tiles) -= av * Ub; meanwhile CN receives the next row's tiles and
        // av[k] is refilled IN PLACE with the next row's left operand as soon as its two DMMA groups have been issued
        // (no second operand buffer: the registers saved keep Ub's eight tiles of a k-step live, so that the sixteen
        // DMMAs of a k-step are independent of the shared-memory loads in flight).
        auto do_row = [&](int i, double2 (&CC)[8], double2 (&CN)[8]) -> bool {
          const int64_t Ir = Is + 8 * i + warp;
          const bool more = i < nwin;
          const bool nxt = more && l_ready(i + 1);
          const double* lnext = tile(Ir + 8, Is) + 2 * lane;
          // Ub tiles of a k-step come from shared memory in two halves of four, one half in flight while the eight
          // DMMAs of the other issue; the next row's accumulator tiles are requested after the first k-step (their
          // scoreboard would otherwise hold back this row's first operand)
          const double* utp = Ut + 2 * lane;
          double2 ua[4], ub[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) ua[c] = *reinterpret_cast<const double2*>(utp + c * 64);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double2 na = neg2(av[k]);
#pragma unroll
            for (int c = 0; c < 4; ++c) ub[c] = *reinterpret_cast<const double2*>(utp + (k * 8 + 4 + c) * 64);
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma884(CC[c].x, CC[c].y, na.x, ua[c].x);
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma884(CC[c].x, CC[c].y, na.y, ua[c].y);
            if (k < 7) {
#pragma unroll
              for (int c = 0; c < 4; ++c) ua[c] = *reinterpret_cast<const double2*>(utp + ((k + 1) * 8 + c) * 64);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma884(CC[4 + c].x, CC[4 + c].y, na.x, ub[c].x);
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma884(CC[4 + c].x, CC[4 + c].y, na.y, ub[c].y);
            if (nxt) av[k] = ldcg_v2(lnext + k * 64);
            if (k == 1 && more) {
              const double* csrc = tile(Ir + 8, Jc) + 2 * lane;
#pragma unroll
              for (int c = 0; c < 8; ++c) CN[c] = *reinterpret_cast<const double2*>(csrc + c * 64);
            }
          }
          if (j == 1 && pending_pub) {   // the stores of the previous row have drained by now: the fence is cheap
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_gpu_u64(f_l, (unsigned long long)((s + 1) * 8 + (pending_pub - 1)));
            pending_pub = 0;
          }
          if (!(i == 1 && j > 1 && a.G == KB)) {   // (that row stays in registers and is overwritten by Ub next step)
            double* cdst = tile(Ir, Jc) + 2 * lane;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(cdst + c * 64) = CC[c];
          }
          if (i == 1) {
            if (j > 1) {   // my column's next pivot row: stays in registers for the next step's U phase
#pragma unroll
              for (int c = 0; c < 8; ++c) unext[c] = CC[c];
              kept = true;
            } else {       // next pivot block D_{s+1}: critical path, published at once
              __threadfence();
              __syncwarp();
              if (lane == 0) st_release_gpu_u64(f_d, (unsigned long long)(s + 1));
            }
          } else if (j == 1) {
            pending_pub = i;
          }
          if (more && !nxt) {
            WT_ADD(4, tmark);
            if (need_flag(i + 1) && !wide_wait_ge(f_l, (unsigned long long)(s * 8 + i + 1), a.abort_word)) return false;
            WT_ADD(5, tmark);
#pragma unroll
            for (int k = 0; k < 8; ++k) av[k] = ldcg_v2(lnext + k * 64);
          }
          return true;
        };
        for (int i = 1; i <= nwin; i += 2) {
          if (!do_row(i, cc, ccn)) return;
          if (i + 1 <= nwin && !do_row(i + 1, ccn, cc)) return;
        }
        if (j == 1 && pending_pub) {
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_gpu_u64(f_l, (unsigned long long)((s + 1) * 8 + (pending_pub - 1)));
        }
      }
      have_u = kept && (a.G == KB);   // one column per CTA and step: the kept row is next step's pivot row
      WT_ADD(4, tmark);
    }
    if (TRACE && group == 0 && threadIdx.x == 0 && a.trace) for (int q = 0; q < 8; ++q) a.trace[role * 8 + q] = tr[q];
  }
}

// --------------------------------------------------------------------------------------------
// factor every partition of `band` (tile-row boundaries d_pstart[0..P], multiples of 8) in place
int spk_wide_lu(spk_ctx* c, double* band, const int64_t* d_pstart, int P) {
  if (P <= 0) return SPK_OK;
  WideLuArgs a;
  a.band = band; a.tpr = c->L.tpr; a.kts = c->L.kt; a.KB = c->kb; a.P = P;
  a.G = c->wide_G > 0 ? std::min(c->wide_G, c->kb) : c->kb;
  a.pstart = d_pstart; a.flags = c->wide_flags; a.abort_word = c->wide_abort;
  a.boost_thr = c->opts.boost_rel * c->anorm_max; a.boost_count = (long long*)c->d_boost;
  a.trace = (band == c->band) ? (long long*)c->lu_trace : nullptr;   // (the band LU only, not the window / reduced ones)
  const void* kern = a.trace ? (const void*)k_wide_lu<true> : (const void*)k_wide_lu<false>;
  if (P > c->wide_flag_parts) { SPK_SET_ERR(c, "wide LU: %d partitions exceed the flag array (%d)", P, c->wide_flag_parts); return SPK_ERR_STATE; }
  const size_t smem = 3 * 4096 * sizeof(double);
  SPK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SPK_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wide_lu<false>, WL_THREADS, smem));
  const int resident = per_sm * c->sm_count;
  const int gsz = a.G + 1;
  int ngroups = std::min(P, resident / gsz);
  if (ngroups < 1) { SPK_SET_ERR(c, "wide LU: a group of %d CTAs is not co-resident on this device", gsz); return SPK_ERR_UNSUPPORTED; }
  SPK_CUDA(c, cudaMemsetAsync(c->wide_flags, 0, sizeof(unsigned long long) * (size_t)P * WIDE_FLAGS_PER_PART, c->stream));
  void* kargs[] = {(void*)&a};
  SPK_CUDA(c, cudaLaunchCooperativeKernel(kern, dim3(ngroups * gsz), dim3(WL_THREADS), kargs, smem, c->stream));
  c->launches++;
  return SPK_OK;
}
