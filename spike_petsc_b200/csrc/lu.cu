// lu.cu -- per-partition banded LU (no pivoting, diagonal boosting) on FP64 tensor cores.
//
// Replaces PCSetUp(inner) = PETSc sparse LU of the AIJ band (/root/reference/src/matbanded.c:178).
//
// One CTA per SPIKE partition, KT+1 warps (KT = ceil(K/8)):
//   * warp w < KT owns tile COLUMN J == w (mod KT) of the sliding KT x KT-tile trailing window; its
//     KT tiles live in registers as DMMA m8n8k4 accumulator fragments (2 doubles/lane/tile), so
//     the whole K x K window is register resident and a tile row of HBM data is touched once.
//   * per 8-pivot step s:  L21(I) = A(I,s) U11^-1 and U12(J) = L11^-1 A(s,J) are two DMMAs per tile
//     (explicit 8x8 inverses), published to shared memory in fragment order; the trailing update
//     A(I,J) -= L21(I) U12(J) is two DMMAs per tile with one 16 B LDS per lane for the A fragment.
//   * the service warp (warp KT) factors the NEXT 8x8 diagonal tile (handed over as soon as its
//     owner has updated it) while the column warps finish the current update, so the sequential
//     pivot chain (~70 cycles/pivot, profiles/r01_microbench_notes.md) is off the critical path;
//     it also feeds a 4-deep ring of cp.async.bulk (TMA) copies that stages the 2KT+1 tiles entering
//     the window two steps ahead.
//   * REV=true runs the same elimination on the row/column-reversed matrix (= UL factorisation of the
//     partition's first tipT tile rows) without storing factors: it only yields the top Schur
//     block S_t needed for the W^(t) spike tip.
// Outputs (FWD): L and U tiles in place (L unit-lower multipliers, U incl. diagonal), the bottom
// Schur block S_b of each partition (for V^(b)), inverse diagonal blocks dinv (for the sweeps),
// boosted-pivot count.
#include "common.cuh"

#define LU_NSTAGE 4

template <int KT>
struct LuSmem {
  double Lfrag[KT][64];   // -L21 tiles, A-fragment order interleaved: [lane*2 + h] = -L[lane/4][4h + lane%4]
  double Ufrag[KT][64];   //  U12 tiles, B-fragment order interleaved: [lane*2 + h] =  U[4h + lane%4][lane/4]
  double Praw[KT][64];    // raw pivot-column tiles A(s+1+i, s), row-major
  double Rraw[KT][64];    // per-warp scratch (C-fragment -> B-fragment conversion)
  double Dtile[64];       // diagonal tile handed to the service warp (row-major)
  double Dlu[64];         // its packed L\U factors
  double Dinv[64];        // packed inverses: strictly lower = Linv, upper incl. diag = Uinv
  double LinvF[64];       // A-fragment order of L11^-1
  double UinvF[64];       // B-fragment order of U11^-1
  double stage[LU_NSTAGE][2 * KT + 1][64];
  unsigned long long full[LU_NSTAGE];
};

struct LuArgs {
  double* band;
  double* dinv;           // nt tiles (FWD only)
  double* schur;          // P * kp*kp : S_b (FWD) or S_t (REV)
  const int64_t* pstart;  // P+1 tile-row boundaries
  long long* boost_count;
  int tpr;
  int tipT;               // REV: window length in tile rows
  int first_part;         // REV: first partition index handled by blockIdx 0
  double boost_thr;
};

template <int KT, bool REV>
__global__ void __launch_bounds__((KT + 1) * 32) k_band_lu(const LuArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LuSmem<KT>& S = *reinterpret_cast<LuSmem<KT>*>(smem_raw);
  constexpr int NT = (KT + 1) * 32;
  constexpr int NCOL = KT * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  auto ld_elem = [&](const double* tile, int idx) -> double { return tile[REV ? 63 - idx : idx]; };
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

  if (warp == KT) {
    // =========================== service warp ===========================================
    auto issue_stage = [&](int s) {
      if (!stage_valid(s)) return;
      uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[s % LU_NSTAGE]);
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)((2 * KT + 1) * 512));
      __syncwarp();
      if (lane == 0) {
        // row chunk: logical tiles (s+KT, s .. s+KT) are contiguous in memory
        const double* src = REV ? tptr(s + KT, s + KT) : tptr(s + KT, s);
        bulk_g2s(S.stage[s % LU_NSTAGE][0], src, (KT + 1) * 512, bar);
      } else if (lane <= KT) {
        const int i = lane - 1;
        bulk_g2s(S.stage[s % LU_NSTAGE][KT + 1 + i], tptr(s + i, s + KT), 512, bar);
      }
    };
    for (int s = 0; s < LU_NSTAGE && s < T; ++s) issue_stage(s);

    for (int s = 0; s < T; ++s) {
      named_bar_sync(2, 64);  // diagonal tile of step s is in S.Dtile
      if (s >= 2 && s + 2 >= LU_NSTAGE) issue_stage(s + 2);
      // ---- 8x8 LU, no pivoting, boosting.  Lane r < 8 holds row r.
      double row[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) row[c] = S.Dtile[(lane & 7) * 8 + c];
      double rinv[8];
      int nboost = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double piv = __shfl_sync(0xffffffffu, row[k], k);
        if (fabs(piv) < a.boost_thr) {
          piv = (piv < 0.0) ? -a.boost_thr : a.boost_thr;
          if (lane == k) row[k] = piv;
          ++nboost;
        }
        rinv[k] = 1.0 / piv;
        double l = row[k] * rinv[k];
#pragma unroll
        for (int c = k + 1; c < 8; ++c) {
          const double u = __shfl_sync(0xffffffffu, row[c], k);
          if (lane > k) row[c] = fma(-l, u, row[c]);
        }
        if (lane > k) row[k] = l;
      }
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) S.Dlu[lane * 8 + c] = row[c];
      }
      if (lane == 0 && nboost) atomicAdd((unsigned long long*)a.boost_count, (unsigned long long)nboost);
      __syncwarp();
      // ---- explicit inverses: lanes 0..7 column c of L11^-1, lanes 8..15 column c of U11^-1
      double x[8];
      const int c = lane & 7;
      if (lane < 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
          for (int j = 0; j < r; ++j) if (j >= c) v = fma(-S.Dlu[r * 8 + j], x[j], v);
          x[r] = (r < c) ? 0.0 : v;
        }
      } else if (lane < 16) {
#pragma unroll
        for (int r = 7; r >= 0; --r) {
          double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
          for (int j = r + 1; j < 8; ++j) if (j <= c) v = fma(-S.Dlu[r * 8 + j], x[j], v);
          x[r] = (r > c) ? 0.0 : v * rinv[r];
        }
      }
      if (lane < 8) {
        // Linv[r][c], A-fragment interleaved index 2*(4r + (c&3)) + (c>>2)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S.LinvF[2 * (4 * r + (c & 3)) + (c >> 2)] = x[r];
          if (r > c) S.Dinv[r * 8 + c] = x[r];
        }
      } else if (lane < 16) {
        // Uinv[r][c] (k=r), B-fragment interleaved index 2*(4c + (r&3)) + (r>>2)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          S.UinvF[2 * (4 * c + (r & 3)) + (r >> 2)] = x[r];
          if (r <= c) S.Dinv[r * 8 + c] = x[r];
        }
      }
      __syncwarp();
      if (!REV) {
        double* dst = tptr(s, s);
        *reinterpret_cast<double2*>(dst + 2 * lane) = *reinterpret_cast<const double2*>(&S.Dlu[2 * lane]);
        double* di = a.dinv + (base + s) * SPK_TILE_ELEMS;
        *reinterpret_cast<double2*>(di + 2 * lane) = *reinterpret_cast<const double2*>(&S.Dinv[2 * lane]);
      }
      __threadfence_block();
      named_bar_arrive(1, NT);  // L11^-1 / U11^-1 of step s are published
    }
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

  // publish the pivot column of step `sn` (this warp's column == sn): diag handed over separately
  auto publish_panel = [&](int sn) {
#pragma unroll
    for (int i = 1; i < KT; ++i) *reinterpret_cast<double2*>(&S.Praw[i - 1][2 * lane]) = acc[i];
    stage_wait(sn);
    const double2 e = stage_valid(sn) ? ld_pair(stage_row(sn, 0)) : make_double2(0.0, 0.0);
    *reinterpret_cast<double2*>(&S.Praw[KT - 1][2 * lane]) = e;
  };
  auto hand_diag = [&](const double2& d) {
    *reinterpret_cast<double2*>(&S.Dtile[2 * lane]) = d;
    __threadfence_block();
    named_bar_arrive(2, 64);
  };

  if (jrel == 0) { hand_diag(acc[0]); publish_panel(0); }

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
    named_bar_sync(1, NT);  // S1: U11^-1/L11^-1(s) ready, Praw(s) published, update(s-1) complete
    // ---------------- phase A: L21 tile `warp` and U12 of the owned column ----------------
    {
      const double* pr = S.Praw[warp];
      const double a0 = pr[g * 8 + tq], a1 = pr[g * 8 + 4 + tq];
      const double2 ub = *reinterpret_cast<const double2*>(&S.UinvF[2 * lane]);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, a0, ub.x);
      dmma884(x0, x1, a1, ub.y);
      if (!REV && s + 1 + warp < T) *reinterpret_cast<double2*>(tptr(s + 1 + warp, s) + 2 * lane) = make_double2(x0, x1);
      const int idx = 2 * (4 * g + 2 * (tq & 1)) + (tq >> 1);
      S.Lfrag[warp][idx] = -x0;
      S.Lfrag[warp][idx + 2] = -x1;
    }
    {
      double b0, b1;
      const int J = s + (jrel == 0 ? KT : jrel);  // column whose U12 tile this warp produces
      if (jrel != 0) {
        double* rr = S.Rraw[warp];
        *reinterpret_cast<double2*>(rr + 2 * lane) = acc[0];
        __syncwarp();
        b0 = rr[tq * 8 + g];
        b1 = rr[(4 + tq) * 8 + g];
      } else {
        stage_wait(s);
        const bool v = stage_valid(s);
        b0 = v ? ld_elem(stage_col(s, 0), tq * 8 + g) : 0.0;
        b1 = v ? ld_elem(stage_col(s, 0), (4 + tq) * 8 + g) : 0.0;
      }
      const double2 la = *reinterpret_cast<const double2*>(&S.LinvF[2 * lane]);
      double y0 = 0.0, y1 = 0.0;
      dmma884(y0, y1, la.x, b0);
      dmma884(y0, y1, la.y, b1);
      if (!REV && J < T) *reinterpret_cast<double2*>(tptr(s, J) + 2 * lane) = make_double2(y0, y1);
      double* uf = S.Ufrag[J - s - 1];
      uf[2 * (8 * tq + (g & 3)) + (g >> 2)] = y0;
      uf[2 * (8 * tq + 4 + (g & 3)) + (g >> 2)] = y1;
    }
    named_bar_sync(3, NCOL);  // S2: Lfrag/Ufrag(s) complete
    // ---------------- phase B: trailing update + window slide ----------------
    stage_wait(s);
    const bool sv = stage_valid(s);
    if (jrel != 0) {
      const double2 bf = *reinterpret_cast<const double2*>(&S.Ufrag[jrel - 1][2 * lane]);
#pragma unroll
      for (int i = 1; i < KT; ++i) {
        const double2 af = *reinterpret_cast<const double2*>(&S.Lfrag[i - 1][2 * lane]);
        dmma884(acc[i].x, acc[i].y, af.x, bf.x);
        dmma884(acc[i].x, acc[i].y, af.y, bf.y);
        if (i == 1 && jrel == 1 && s + 1 < T) hand_diag(acc[1]);  // next diagonal tile: to the service warp now
      }
      double2 f = sv ? ld_pair(stage_row(s, jrel)) : make_double2(0.0, 0.0);
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
      const double2 bf = *reinterpret_cast<const double2*>(&S.Ufrag[KT - 1][2 * lane]);
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
    if (jrel == 0 && s + 1 < T) publish_panel(s + 1);
  }
}

// --------------------------------------------------------------------------------------------
template <int KT, bool REV>
static int launch_lu_kt(spk_ctx* c, int grid, int first_part) {
  LuArgs a;
  a.band = c->band; a.dinv = c->dinv; a.schur = REV ? c->St : c->Sb; a.pstart = c->d_pstart;
  a.boost_count = (long long*)c->d_boost; a.tpr = c->L.tpr; a.tipT = c->tipT; a.first_part = first_part;
  a.boost_thr = c->opts.boost_rel * c->anorm_max;
  const size_t smem = sizeof(LuSmem<KT>);
  SPK_CUDA(c, cudaFuncSetAttribute(k_band_lu<KT, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_band_lu<KT, REV><<<grid, (KT + 1) * 32, smem, c->stream>>>(a);
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

// UL windows for W^(t): partitions 1..P-1 (and partition 0 when a left-neighbour rank exists)
int spk_launch_ul_tips(spk_ctx* c) {
  const int first = (c->opts.rank > 0) ? 0 : 1;
  const int grid = c->P - first;
  if (grid <= 0) return SPK_OK;
  return launch_lu<true>(c, grid, first);
}
