// lu.cu -- per-partition banded block-LU (no pivoting, diagonal boosting) on FP64 tensor cores.
//
// Replaces PCSetUp(inner) = PETSc sparse LU of the AIJ band (/root/reference/src/matbanded.c:178).
//
// Factorisation computed (8x8 tiles, tile row/col indices I,J):   A = Lb * Ub  with
//     Lb(I,I) = I,  Lb(I,J) = A~(I,J) * D_J^-1  (J < I),      Ub(I,J) = A~(I,J)  (J >= I),  D_I = A~(I,I)
// where A~ is the Schur-updated matrix.  The band keeps Lb tiles below the diagonal, Ub tiles above
// it and the EXPLICIT INVERSE D_I^-1 in the diagonal tile slot (that is what both this kernel and
// the triangular sweeps need; a scalar LU of D_I never has to be applied).  Mathematically this is
// the same elimination as the scalar no-pivot LU (same Schur complements, same pivots, same
// boosting rule), grouped by 8 pivots.
//
// One CTA per SPIKE partition, KT+1 warps (KT = ceil(K/8)):
//   * warp w < KT owns tile COLUMN J == w (mod KT) of the sliding KT x KT-tile trailing window; its
//     KT tiles live in registers as DMMA m8n8k4 accumulator fragments (2 doubles/lane/tile), so the
//     whole K x K window is register resident and every band entry is read and written once.
//   * per 8-pivot step s: warp w forms Lb(s+1+w, s) = A~(s+1+w, s) D_s^-1 (2 DMMAs, fragments
//     published to shared memory), then every warp updates its column: A~(I,J) -= Lb(I,s) A~(s,J)
//     (2 DMMAs per tile, one 16 B LDS per lane for the A fragment; the B fragment is the raw pivot-row
//     tile its owner published at the end of the previous step).
//   * the service warp inverts the NEXT diagonal tile by in-register Gauss-Jordan (lane r = row r,
//     pivot row broadcast by shuffles) as soon as its owner has updated it, i.e. concurrently with
//     the current trailing update; it also keeps a 4-deep ring of cp.async.bulk (TMA) copies filled
//     with the 2KT+1 tiles that enter the window two steps ahead.
//   * REV=true runs the same elimination on the row/column-reversed matrix (= bottom-up elimination
//     of the partition's first tipT tile rows) without storing factors: it only yields the top
//     Schur block S_t needed for the W^(t) spike tip.
// Outputs (FWD): block factors in place, the bottom Schur block S_b of each partition (for V^(b)),
// boosted-pivot count.
#include "common.cuh"

#define LU_NSTAGE 4
#define LU_TRACE_STEPS 64
// optional clock64 trace of CTA 0 (tools/lu_trace.py): [step-100][16]; 0..7 column warp 0, 8..15 service
// LU_TRB: stamp after a barrier -- BAR.SYNC is deferred-blocking, so first consume a shared-memory word
#define LU_TRB(slot, ptr) do { if (a.trace && blockIdx.x == 0 && s >= 100 && s < 100 + LU_TRACE_STEPS) { double d_ = *(volatile double*)(ptr); double e_; asm volatile("add.f64 %0, %1, %1;" : "=d"(e_) : "d"(d_)); if (lane == 0) a.trace[(s - 100) * 16 + (slot)] = clock64() + (e_ == 1.2345e300 ? 1 : 0); } } while (0)
#define LU_TR(slot) do { if (a.trace && blockIdx.x == 0 && lane == 0 && s >= 100 && s < 100 + LU_TRACE_STEPS) a.trace[(s - 100) * 16 + (slot)] = clock64(); } while (0)

// Warp placement.  FP64 instructions of a warp queue IN ORDER behind every DMMA already issued on the
// same SM sub-partition (hardware warp id % 4): with the column warps streaming 2*KT DMMAs each, an
// FP64 op of the service warp would wait for the whole queue (>1000 cycles, tools/lu_trace.py) and
// the Gauss-Jordan chain could never overlap the trailing update.  With LU_DEDICATED the service warp
// is hardware warp 3 and every other warp id == 3 (mod 4) is parked, so sub-partition 3 carries no
// DMMA stream.
#ifndef LU_DEDICATED
#define LU_DEDICATED 0
#endif
template <int KT>
constexpr int lu_total_warps() {
  if (!LU_DEDICATED) return KT + 1;
  int W = 4;
  while (W - W / 4 < KT) ++W;
  return W;
}

template <int KT>
struct LuSmem {
  double Lfrag[KT][64];      // -Lb(s+1+i, s), A-fragment order interleaved: [lane*2 + h] = -L[lane/4][4h + lane%4]
  double Ufrag[2][KT][64];   //  A~(s, s+1+j), B-fragment order interleaved: [lane*2 + h] =  U[4h + lane%4][lane/4]
  double Praw[KT][64];       // raw pivot-column tiles A~(s+1+i, s), row-major
  double Dtile[64];          // diagonal tile handed to the service warp (row-major)
  double Dinv[64];           // its inverse, row-major (goes to the band's diagonal slot)
  double DinvF[64];          // its inverse in B-fragment order
  double stage[LU_NSTAGE][2 * KT + 1][64];
  unsigned long long full[LU_NSTAGE];
};

struct LuArgs {
  double* band;
  double* schur;          // P * kp*kp : S_b (FWD) or S_t (REV)
  const int64_t* pstart;  // P+1 tile-row boundaries
  long long* boost_count;
  int tpr;
  int tipT;               // REV: window length in tile rows
  int first_part;         // REV: first partition index handled by blockIdx 0
  double boost_thr;
  long long* trace;       // optional debug stamps; nullptr in production
};

template <int KT, bool REV>
__global__ void __launch_bounds__(lu_total_warps<KT>() * 32, (LU_DEDICATED ? 1 : (KT >= 12 ? 2 : (KT >= 8 ? 3 : 4)))) k_band_lu(const LuArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LuSmem<KT>& S = *reinterpret_cast<LuSmem<KT>*>(smem_raw);
  constexpr int NT = (KT + 1) * 32;
  constexpr int NCOL = KT * 32;
  const int hw = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_service = LU_DEDICATED ? (hw == 3) : (hw == KT);
  const int warp = LU_DEDICATED ? hw - ((hw + 1) >> 2) : hw;  // column index of a column warp
  const bool parked = LU_DEDICATED && !is_service && (((hw & 3) == 3) || warp >= KT);
  const int g = lane >> 2, tq = lane & 3;
  const int part = blockIdx.x + (REV ? a.first_part : 0);
  const int64_t t0 = a.pstart[part];
  const int64_t plen = a.pstart[part + 1] - t0;
  const int T = REV ? (int)(plen < a.tipT ? plen : a.tipT) : (int)plen;
  const int64_t base = REV ? (t0 + T - 1) : t0;  // actual tile index of logical tile 0
  const int tpr = a.tpr;

  // logical (I,J) -> address of the tile in the band
  auto tptr = [&](int I, int J) -> double* {
    const int64_t Ia = REV ? base - I : base + I;
    const int64_t Ja = REV ? base - J : base + J;
    return a.band + (Ia * tpr + (Ja - Ia + KT)) * SPK_TILE_ELEMS;
  };
  // logical C-fragment pair (elements 2*lane, 2*lane+1) of a row-major tile stored in ACTUAL orientation
  auto ld_pair = [&](const double* tile) -> double2 {
    if (!REV) return *reinterpret_cast<const double2*>(tile + 2 * lane);
    const double2 v = *reinterpret_cast<const double2*>(tile + 62 - 2 * lane);
    return make_double2(v.y, v.x);
  };
  auto stage_valid = [&](int s) -> bool { return s + KT < T; };
  auto stage_wait = [&](int s) {
    if (stage_valid(s)) mbar_wait(reinterpret_cast<uint64_t*>(&S.full[s % LU_NSTAGE]), (uint32_t)((s / LU_NSTAGE) & 1));
  };
  // staged tile of the row chunk (s+KT, s+j), j = 0..KT   /   of the column (s+i, s+KT), i = 0..KT-1
  auto stage_row = [&](int s, int j) -> const double* { return S.stage[s % LU_NSTAGE][REV ? KT - j : j]; };
  auto stage_col = [&](int s, int i) -> const double* { return S.stage[s % LU_NSTAGE][KT + 1 + i]; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < LU_NSTAGE; ++i) mbar_init(reinterpret_cast<uint64_t*>(&S.full[i]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (parked) return;

  if (is_service) {
    // =========================== service warp ===========================================
    // Staging of step s: the service warp arms the mbarrier (expect_tx for all 2KT+1 tiles) and copies
    // the contiguous row chunk; the KT single-tile column copies are issued one per column warp
    // (a per-lane UBLKCP would be serialised through the uniform datapath, ~80 cycles each).
    auto issue_stage = [&](int s, bool with_columns) {
      if (!stage_valid(s)) return;
      uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[s % LU_NSTAGE]);
      if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)((2 * KT + 1) * 512));
        // row chunk: logical tiles (s+KT, s .. s+KT) are contiguous in memory
        const double* src = REV ? tptr(s + KT, s + KT) : tptr(s + KT, s);
        bulk_g2s(S.stage[s % LU_NSTAGE][0], src, (KT + 1) * 512, bar);
      }
      if (with_columns) {
        __syncwarp();
        if (lane >= 1 && lane <= KT) bulk_g2s(S.stage[s % LU_NSTAGE][KT + lane], tptr(s + lane - 1, s + KT), 512, bar);
      }
    };
    // prologue: steps 0 and 1 completely; steps 2,3 get their column tiles from the column warps
    for (int s = 0; s < LU_NSTAGE && s < T; ++s) issue_stage(s, s < 2);
    const int r8 = lane & 7;
    int nboost_total = 0;

    for (int s = 0; s < T; ++s) {
      LU_TR(8);
      named_bar_sync(2, 64);  // diagonal tile of step s is in S.Dtile
      LU_TRB(9, &S.Dtile[0]);
      // ---- in-place Gauss-Jordan inverse of the 8x8 pivot block, no pivoting, boosting.
      //      Lane r (mod 8) holds row r.  Per pivot the dependent chain is
      //      shuffle(pivot) -> reciprocal -> multiplier -> one FMA (the next pivot element).
      double row[8];
      {
        const double2* dt = reinterpret_cast<const double2*>(&S.Dtile[r8 * 8]);
        const double2 q0 = dt[0], q1 = dt[1], q2 = dt[2], q3 = dt[3];
        row[0] = q0.x; row[1] = q0.y; row[2] = q1.x; row[3] = q1.y;
        row[4] = q2.x; row[5] = q2.y; row[6] = q3.x; row[7] = q3.y;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double piv = __shfl_sync(0xffffffffu, row[k], k);
        // reciprocal: hardware seed r0 (~2^-20) + one Halley step, 1/piv = r0 (1 + e + e^2), e = 1 - piv r0,
        // relative error e^3.  The multiplier is formed from q = a*r0 in parallel: f = q + q (e + e^2).
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(piv));
        const bool isp = (r8 == k);
        const double q = isp ? 0.0 : row[k] * r0;
        const double e = fma(-piv, r0, 1.0);
        const double t = fma(e, e, e);
        double f = fma(q, t, q);             // multiplier of the pivot row for this lane's row (0 for the pivot row)
        double rc = fma(r0, t, r0);          // 1/piv to fp64 accuracy
        if (fabs(piv) < a.boost_thr) {       // warp-uniform, rare: boosted pivot (SpikeGPU-style)
          rc = (piv < 0.0) ? -1.0 / a.boost_thr : 1.0 / a.boost_thr;
          f = isp ? 0.0 : row[k] * rc;
          ++nboost_total;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c == k) continue;
          const double u = __shfl_sync(0xffffffffu, row[c], k);   // unscaled pivot-row entry
          row[c] = isp ? u * rc : fma(-f, u, row[c]);
        }
        row[k] = isp ? rc : -f;              // column k of the inverse-in-progress
      }
      if (lane < 8) {
        double2* dl = reinterpret_cast<double2*>(&S.Dinv[lane * 8]);
        dl[0] = make_double2(row[0], row[1]); dl[1] = make_double2(row[2], row[3]);
        dl[2] = make_double2(row[4], row[5]); dl[3] = make_double2(row[6], row[7]);
        // B-fragment order: element Dinv[k=r][c] -> [2*(4c + (r&3)) + (r>>2)]
        double* f0 = &S.DinvF[2 * (lane & 3) + (lane >> 2)];
#pragma unroll
        for (int c = 0; c < 8; ++c) f0[8 * c] = row[c];
      }
      __syncwarp();
      LU_TR(11);
      named_bar_arrive(1, NT);  // D_s^-1 is published
      LU_TR(12);
      if (!REV) {  // factor output, off the critical path
        double* dst = tptr(s, s);
        *reinterpret_cast<double2*>(dst + 2 * lane) = *reinterpret_cast<const double2*>(&S.Dinv[2 * lane]);
      }
      if (s >= 2) issue_stage(s + 2, false);
    }
    if (lane == 0 && nboost_total) atomicAdd((unsigned long long*)a.boost_count, (unsigned long long)nboost_total);
    return;
  }

  // =========================== column warps ================================================
  double2 acc[KT];
  // initial window: rows 0..KT-1 of column `warp`
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    acc[i] = (i < T && warp < T) ? ld_pair(tptr(i, warp)) : make_double2(0.0, 0.0);
  }
  int jrel = warp;  // (column owned) - s

  // B-fragment publication of a C-fragment tile: element (k=g, c=2tq+e) -> [2*(4c + (g&3)) + (g>>2)]
  auto publish_u = [&](double* uf, const double2& v) {
    uf[2 * (8 * tq + (g & 3)) + (g >> 2)] = v.x;
    uf[2 * (8 * tq + 4 + (g & 3)) + (g >> 2)] = v.y;
  };
  // Work done by a warp when step sn is about to start and its column is (sn + jr):
  //   jr == 0 : its column is the pivot column -> publish the raw panel tiles (diag handed separately)
  //             and the band-edge row tile A(sn, sn+KT) as the last U fragment
  //   jr >= 1 : its slot-0 tile is the pivot-row tile A~(sn, sn+jr): it is final -> store it as the
  //             Ub factor and publish it as a B fragment
  auto publish_for_step = [&](int sn, int jr) {
    double* ub = &S.Ufrag[sn & 1][0][0];
    if (jr == 0) {
#pragma unroll
      for (int i = 1; i < KT; ++i) *reinterpret_cast<double2*>(&S.Praw[i - 1][2 * lane]) = acc[i];
      stage_wait(sn);
      const bool v = stage_valid(sn);
      const double2 e = v ? ld_pair(stage_row(sn, 0)) : make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(&S.Praw[KT - 1][2 * lane]) = e;
      const double2 ue = v ? ld_pair(stage_col(sn, 0)) : make_double2(0.0, 0.0);
      publish_u(ub + (KT - 1) * 64, ue);
    } else {
      publish_u(ub + (jr - 1) * 64, acc[0]);
      if (!REV && sn + jr < T) *reinterpret_cast<double2*>(tptr(sn, sn + jr) + 2 * lane) = acc[0];
    }
  };
  auto hand_diag = [&](const double2& d) {
    *reinterpret_cast<double2*>(&S.Dtile[2 * lane]) = d;
    __syncwarp();
    named_bar_arrive(2, 64);  // barrier arrival orders the shared-memory writes for the waiting service warp
  };

  if (jrel == 0) hand_diag(acc[0]);
  publish_for_step(0, jrel);

  const int kp = KT * 8;
  for (int s = 0; s < T; ++s) {
    if (s == T - KT) {
      // registers hold the trailing Schur complement of the partition (rows/cols s..s+KT-1)
      double* out = a.schur + (int64_t)part * kp * kp;
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        const int r = 8 * i + g, cc = 8 * jrel + 2 * tq;
        if (!REV) {
          *reinterpret_cast<double2*>(out + (int64_t)r * kp + cc) = acc[i];
        } else {
          out[(int64_t)(kp - 1 - r) * kp + (kp - 1 - cc)] = acc[i].x;
          out[(int64_t)(kp - 1 - r) * kp + (kp - 2 - cc)] = acc[i].y;
        }
      }
    }
    if (warp == 0) LU_TR(0);
    named_bar_sync(1, NT);  // S1: D_s^-1 ready, Praw/Ufrag(s) published, update(s-1) complete
    if (warp == 0) LU_TRB(1, &S.DinvF[0]);
    // ---------------- phase A: Lb(s+1+warp, s) = A~(s+1+warp, s) * D_s^-1 ----------------
    {
      const double* pr = S.Praw[warp];
      const double a0 = pr[g * 8 + tq], a1 = pr[g * 8 + 4 + tq];
      const double2 ub = *reinterpret_cast<const double2*>(&S.DinvF[2 * lane]);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, a0, ub.x);
      dmma884(x0, x1, a1, ub.y);
      if (!REV && s + 1 + warp < T) *reinterpret_cast<double2*>(tptr(s + 1 + warp, s) + 2 * lane) = make_double2(x0, x1);
      const int idx = 2 * (4 * g + 2 * (tq & 1)) + (tq >> 1);
      S.Lfrag[warp][idx] = -x0;
      S.Lfrag[warp][idx + 2] = -x1;
    }
    if (warp == 0) LU_TR(2);
    named_bar_sync(3, NCOL);  // S2: Lfrag(s) complete
    if (warp == 0) LU_TRB(3, &S.Lfrag[0][0]);
    // column tile `warp` of step s+2 (its ring slot was last read in update(s-2), which is complete)
    if (lane == 0 && stage_valid(s + 2)) {
      const int sn = s + 2;
      bulk_g2s(S.stage[sn % LU_NSTAGE][KT + 1 + warp], tptr(sn + warp, sn + KT), 512,
               reinterpret_cast<uint64_t*>(&S.full[sn % LU_NSTAGE]));
    }
    // ---------------- phase B: trailing update + window slide ----------------
    const double* ufs = &S.Ufrag[s & 1][0][0];
    if (jrel != 0) {
      const double2 bf = *reinterpret_cast<const double2*>(ufs + (jrel - 1) * 64 + 2 * lane);
#pragma unroll
      for (int i = 1; i < KT; ++i) {
        const double2 af = *reinterpret_cast<const double2*>(&S.Lfrag[i - 1][2 * lane]);
        dmma884(acc[i].x, acc[i].y, af.x, bf.x);
        dmma884(acc[i].x, acc[i].y, af.y, bf.y);
        if (i == 1 && jrel == 1 && s + 1 < T) hand_diag(acc[1]);  // next diagonal tile: to the service warp now
      }
      stage_wait(s);
      double2 f = stage_valid(s) ? ld_pair(stage_row(s, jrel)) : make_double2(0.0, 0.0);
      {
        const double2 af = *reinterpret_cast<const double2*>(&S.Lfrag[KT - 1][2 * lane]);
        dmma884(f.x, f.y, af.x, bf.x);
        dmma884(f.x, f.y, af.y, bf.y);
      }
#pragma unroll
      for (int i = 1; i < KT; ++i) acc[i - 1] = acc[i];
      acc[KT - 1] = f;
      --jrel;
    } else {
      // pivot warp: its column is retired; take over the entering column s+KT (all tiles fresh)
      const double2 bf = *reinterpret_cast<const double2*>(ufs + (KT - 1) * 64 + 2 * lane);
      stage_wait(s);
      const bool sv = stage_valid(s);
#pragma unroll
      for (int i = 0; i < KT; ++i) {
        double2 f;
        if (i < KT - 1) f = sv ? ld_pair(stage_col(s, i + 1)) : make_double2(0.0, 0.0);
        else            f = sv ? ld_pair(stage_row(s, KT)) : make_double2(0.0, 0.0);
        const double2 af = *reinterpret_cast<const double2*>(&S.Lfrag[i][2 * lane]);
        dmma884(f.x, f.y, af.x, bf.x);
        dmma884(f.x, f.y, af.y, bf.y);
        acc[i] = f;
      }
      jrel = KT - 1;
    }
    if (warp == 0) LU_TR(4);
    if (s + 1 < T) publish_for_step(s + 1, jrel);
    if (warp == 0) LU_TR(5);
  }
}

// --------------------------------------------------------------------------------------------
template <int KT, bool REV>
static int launch_lu_kt(spk_ctx* c, int grid, int first_part) {
  LuArgs a;
  a.band = c->band; a.schur = REV ? c->St : c->Sb; a.pstart = c->d_pstart;
  a.boost_count = (long long*)c->d_boost; a.tpr = c->L.tpr; a.tipT = c->tipT; a.first_part = first_part;
  a.boost_thr = c->opts.boost_rel * c->anorm_max;
  a.trace = (long long*)c->lu_trace;
  const size_t smem = sizeof(LuSmem<KT>);
  SPK_CUDA(c, cudaFuncSetAttribute(k_band_lu<KT, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_band_lu<KT, REV><<<grid, lu_total_warps<KT>() * 32, smem, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

template <bool REV>
static int launch_lu(spk_ctx* c, int grid, int first_part) {
  switch (c->L.kt) {
#define CASE(K_) case K_: return launch_lu_kt<K_, REV>(c, grid, first_part);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default:
      SPK_SET_ERR(c, "half-bandwidth %d (kt=%d) outside the supported range 1..%d", c->L.k, c->L.kt, 8 * SPK_MAX_KT);
      return SPK_ERR_UNSUPPORTED;
  }
}

int spk_launch_lu(spk_ctx* c) { return launch_lu<false>(c, c->P, 0); }

// bottom-up windows for W^(t): partitions 1..P-1 (and partition 0 when a left-neighbour rank exists)
int spk_launch_ul_tips(spk_ctx* c) {
  const int first = (c->opts.rank > 0) ? 0 : 1;
  const int grid = c->P - first;
  if (grid <= 0) return SPK_OK;
  return launch_lu<true>(c, grid, first);
}
