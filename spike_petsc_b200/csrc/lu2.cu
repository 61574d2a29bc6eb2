// lu2.cu -- per-partition banded block-LU for wide bands (KT >= 8): the kernel of lu.cu with TWO window columns
// per warp.
//
// Same factorisation, same storage, same pipeline as k_band_lu (see the header of lu.cu: register-resident
// sliding window held transposed, packages of pivot-column tiles published tile by tile through shared memory
// with one mbarrier per tile, lookahead warp for the chain of pivot-block inverses, Newton-Schulz inverses on
// the tensor cores).  What changes is the work assignment.  ncu on k_band_lu at KT = 13 shows the shared-memory
// data pipe at 95 %: every column warp reads every package tile (13 x 512 B per step and warp) to feed two DMMAs
// per tile.  Here a warp owns two adjacent columns, so one 16 B shared-memory load per lane feeds FOUR DMMAs
// (two independent accumulation chains per column are in flight), the package traffic halves and the CTA
// shrinks to ceil(KT/2)+1 warps with 128 registers per thread (two CTAs per SM up to KT = 14).
//
// Column slots: VS = 2*ceil(KT/2) virtual slots, slot v holds column c == v (mod VS); at step s slot distance
// d = c - s.  d = 1: the next pivot column (publishes package(s+1) while it is updated); d = 2: hands the
// lookahead warp the tiles of the next-but-one pivot block; d = KT: the entering column.  For odd KT there is
// one slot more than live columns: the slot whose column was just retired idles for one step (d = 0) and uses
// it to load the column that enters at the next step; for even KT the retired slot reloads at once.
#include "lu_dev.cuh"

template <int KT>
struct Lu2Cfg {
  static constexpr int NW = (KT + 1) / 2;      // column warps
  static constexpr int VS = 2 * NW;            // virtual column slots
  static constexpr bool IDLE = (VS > KT);      // one slot idles per step
  static constexpr int NSM = 2;                // newest rows of every column kept in shared memory
  static constexpr int NR = KT - NSM;          // rows of every column kept in registers
};

template <int KT>
struct Lu2Smem {
  double PK[LU_R][KT][64];   // package of step s (slot s % LU_R): tiles 0..KT-2 = -A~(s+1+j, s) row-major, tile KT-1 raw band edge
  double XC[LU_R][64];       // D_s^-1, row-major
  double LT[Lu2Cfg<KT>::VS][Lu2Cfg<KT>::NSM][64];   // per column slot: its newest rows (transposed C-fragment order)
  double tP[2][64];          // handed over during update(u), buffer u & 1, row-major: A~(u+2, u+1)
  double tUt[2][64];         //                                                        A~(u+1, u+2)^T
  double tA[2][64];          //                                                        A~(u+2, u+2)
  double tAt[2][64];         //                                                        A~(u+2, u+2)^T
  unsigned long long xfull[LU_R];       // 1 arrival: the lookahead warp has published D_s^-1
  unsigned long long tfull[LU_R][KT];   // 1 arrival each: package tile j of the slot has been published
  unsigned long long empty[LU_R];       // NW arrivals: every column warp is done with the slot
  unsigned long long tiles[2];          // 2 arrivals: the d = 1 and the d = 2 column have handed their tiles over
};

template <int KT, bool REV>
__global__ void __launch_bounds__((Lu2Cfg<KT>::NW + 1) * 32, (KT <= 14 ? 2 : 1)) k_band_lu2(const LuArgs a) {
  using C = Lu2Cfg<KT>;
  constexpr int NW = C::NW, VS = C::VS, NSM = C::NSM, NR = C::NR;
  constexpr bool IDLE = C::IDLE;
  constexpr bool TRACE = false;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Lu2Smem<KT>& S = *reinterpret_cast<Lu2Smem<KT>*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool is_lookahead = (warp == NW);
  const int g = lane >> 2, tq = lane & 3;
  const int part = blockIdx.x + (REV ? a.first_part : 0);
  const int64_t t0 = a.pstart[part];
  const int64_t plen = a.pstart[part + 1] - t0;
  const int T = REV ? (int)(plen < a.tipT ? plen : a.tipT) : (int)plen;
  const int64_t base = REV ? (t0 + T - 1) : t0;  // actual tile index of logical tile 0
  const int tpr = a.tpr;

  auto tptr = [&](int I, int J) -> double* {
    const int64_t Ia = REV ? base - I : base + I;
    const int64_t Ja = REV ? base - J : base + J;
    return a.band + (Ia * tpr + (Ja - Ia + KT)) * SPK_TILE_ELEMS;
  };
  auto ld_pair = [&](const double* tile) -> double2 {
    if (!REV) return *reinterpret_cast<const double2*>(tile + 2 * lane);
    const double2 v = *reinterpret_cast<const double2*>(tile + 62 - 2 * lane);
    return make_double2(v.y, v.x);
  };
  auto ld_tile = [&](int I, int J) -> double2 {
    return (I < T && J < T) ? ld_pair(tptr(I, J)) : make_double2(0.0, 0.0);
  };
  auto xfull_bar = [&](int s) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.xfull[s % LU_R]); };
  auto tfull_bar = [&](int s, int j) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.tfull[s % LU_R][j]); };
  auto empty_bar = [&](int s) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.empty[s % LU_R]); };
  auto wait_slot_free = [&](int s) {
    if (s >= LU_R) mbar_wait(empty_bar(s), (uint32_t)((s / LU_R - 1) & 1));
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < LU_R; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(&S.xfull[i]), 1);
      for (int j = 0; j < KT; ++j) mbar_init(reinterpret_cast<uint64_t*>(&S.tfull[i][j]), 1);
      mbar_init(reinterpret_cast<uint64_t*>(&S.empty[i]), NW);
    }
    mbar_init(reinterpret_cast<uint64_t*>(&S.tiles[0]), 2);
    mbar_init(reinterpret_cast<uint64_t*>(&S.tiles[1]), 2);
    fence_mbar_init();
  }
  __syncthreads();

  if (is_lookahead) {
    // =========================== lookahead warp (as in lu.cu) ==============================
    const double thr = a.boost_thr, rthr = 1.0 / a.boost_thr;
    int nboost = 0;
    double2 x = make_double2(0.0, 0.0);   // D_{s-1}^-1
    const double2 p1 = ld_tile(1, 0), a1 = ld_tile(1, 1);
    const double2 ut1 = cfrag_transpose(ld_tile(0, 1), g, tq), at1 = cfrag_transpose(a1, g, tq);
    for (int s = 0; s < T; ++s) {
      double2 d, dt;
      if (s == 0) {
        d = ld_pair(tptr(0, 0));
        dt = cfrag_transpose(d, g, tq);
      } else {
        double2 pc, uct;
        if (s == 1) {
          pc = p1; uct = ut1; d = a1; dt = at1;
        } else {        // final after update(s-2): handed over by their owners early in that update
          mbar_wait(reinterpret_cast<uint64_t*>(&S.tiles[s & 1]), (uint32_t)(((s - 2) >> 1) & 1));
          pc = *reinterpret_cast<const double2*>(&S.tP[s & 1][2 * lane]);
          uct = *reinterpret_cast<const double2*>(&S.tUt[s & 1][2 * lane]);
          d = *reinterpret_cast<const double2*>(&S.tA[s & 1][2 * lane]);
          dt = *reinterpret_cast<const double2*>(&S.tAt[s & 1][2 * lane]);
        }
        // Ub^T = U^T X^T ;  D = A - P Ub ;  D^T = A^T - Ub^T P^T     (X = D_{s-1}^-1)
        double2 ubt = make_double2(0.0, 0.0);
        dmma_cc(ubt, uct, x);
        dmma_cc(d, neg2(pc), ubt);
        dmma_cc(dt, neg2(ubt), pc);
      }
      double2 xt;
      jacobi_start8(d, dt, x, xt, g, tq);
      if (!ns_refine8(d, dt, x, xt, g, tq)) {
        const float2 xf = gj8_f32_cfrag(d, g, tq);
        const float2 xft = cfrag_transpose_f(xf, g, tq);
        x = make_double2(f2d_bits(xf.x), f2d_bits(xf.y));
        xt = make_double2(f2d_bits(xft.x), f2d_bits(xft.y));
        if (!ns_refine8(d, dt, x, xt, g, tq)) x = gj8_cfrag(d, g, tq, thr, rthr, nboost);
      }
      wait_slot_free(s);
      *reinterpret_cast<double2*>(&S.XC[s % LU_R][2 * lane]) = x;
      __syncwarp();
      if (lane == 0) mbar_arrive(xfull_bar(s));   // D_s^-1 is published
      if (!REV) *reinterpret_cast<double2*>(tptr(s, s) + 2 * lane) = x;  // factor output
    }
    if (lane == 0 && nboost) atomicAdd((unsigned long long*)a.boost_count, (unsigned long long)nboost);
    return;
  }

  // =========================== column warps: slots A (v = 2*warp) and B (v = 2*warp+1) =============
  constexpr int SGN = REV ? -1 : 1;
  const int RS = SGN * (tpr - 1) * SPK_TILE_ELEMS;   // one tile row down, same column (doubles)
  constexpr int CS = SGN * SPK_TILE_ELEMS;           // one tile column to the right
  const int o0 = REV ? 63 - (16 * tq + g) : 16 * tq + g;
  const int o1 = REV ? o0 - 8 : o0 + 8;
  const int l0 = 16 * tq + g;
  auto ldT = [&](const double* tile) -> double2 { return make_double2(tile[o0], tile[o1]); };
  auto stT = [&](double* tile, const double2& v) { tile[o0] = v.x; tile[o1] = v.y; };
  auto stT_s = [&](double* tile, const double2& v) { tile[l0] = v.x; tile[l0 + 8] = v.y; };
  auto col_tile = [&](int i, int J) -> double2 { return (i < T && J < T) ? ldT(tptr(i, J)) : make_double2(0.0, 0.0); };

  double2 accA[NR], accB[NR];
  double* const ltA = &S.LT[2 * warp][0][2 * lane];       // tail row j at + j*64
  double* const ltB = &S.LT[2 * warp + 1][0][2 * lane];
  int dA = 2 * warp, dB = 2 * warp + 1;                  // column - s; 0 = retired (idle slot, odd KT only)
#pragma unroll
  for (int i = 0; i < NR; ++i) { accA[i] = col_tile(i, dA); accB[i] = col_tile(i, dB); }
#pragma unroll
  for (int j = 0; j < NSM; ++j) {
    *reinterpret_cast<double2*>(ltA + j * 64) = col_tile(NR + j, dA);
    *reinterpret_cast<double2*>(ltB + j * 64) = col_tile(NR + j, dB);
  }
  double* pfA = tptr(KT, dA);    // running pointers: tile (s+KT, c)
  double* pfB = tptr(KT, dB);

  auto stage_edge = [&](int sn, const double* src) {
    double* dst = &S.PK[sn % LU_R][KT - 1][2 * lane];
    if (sn + KT < T) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + 2 * lane) : "memory");
    else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
  };
  // fresh tiles of an entering column; p0 = tile (sn+KT, sn+KT), cv = the column exists
  auto reload = [&](double2 (&acc)[NR], double* lt, const double* p0, bool cv) {
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] = cv ? ldT(p0 - (KT - i) * RS) : make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < NSM; ++j) *reinterpret_cast<double2*>(lt + j * 64) = cv ? ldT(p0 - (NSM - j) * RS) : make_double2(0.0, 0.0);
  };
  auto prefetch_tiles = [&](const double* p, int step, int cnt) {
    for (int l = lane; l < 4 * cnt; l += 32) {
      const double* q = p + (l >> 2) * step + (l & 3) * 16;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
  };
  const int kp = KT * 8;
  auto schur_col = [&](const double2 (&acc)[NR], const double* lt, int jcol) {
    double* out = a.schur + (int64_t)part * kp * kp;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const double2 t = (i < NR) ? acc[i < NR ? i : 0] : *reinterpret_cast<const double2*>(lt + (i < NR ? 0 : i - NR) * 64);
      const int r = 8 * i + 2 * tq, cc = 8 * jcol + g;
      if (!REV) {
        out[(int64_t)r * kp + cc] = t.x;
        out[(int64_t)(r + 1) * kp + cc] = t.y;
      } else {
        out[(int64_t)(kp - 1 - r) * kp + (kp - 1 - cc)] = t.x;
        out[(int64_t)(kp - 2 - r) * kp + (kp - 1 - cc)] = t.y;
      }
    }
  };
  if (T == KT) {
    if (dA < KT) schur_col(accA, ltA, dA);
    if (dB < KT) schur_col(accB, ltB, dB);
  }
  if (dA == 0) {   // warp 0: nothing is eliminated yet, column 0 is published as loaded
    stage_edge(0, pfA);
    asm volatile("cp.async.commit_group;" ::: "memory");
    double* pk0 = &S.PK[0][0][0];
#pragma unroll
    for (int i = 1; i < KT; ++i) {
      const double2 t = (i < NR) ? accA[i < NR ? i : 0] : *reinterpret_cast<const double2*>(ltA + (i < NR ? 0 : i - NR) * 64);
      stT_s(pk0 + (i - 1) * 64, neg2(t));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (lane < KT) mbar_arrive(tfull_bar(0, lane));
    if (!IDLE) {
      pfA += KT * CS;
      reload(accA, ltA, pfA, KT < T);
      dA = KT;
    }
  }

  for (int s = 0; s < T; ++s) {
    const int slot = s % LU_R;
    const uint32_t par = (uint32_t)((s / LU_R) & 1);
    // odd KT: the slot whose column retired idles this step and loads the column that enters at step s+1
    if (IDLE && dA == 0) { pfA += VS * CS; reload(accA, ltA, pfA + RS, s + 1 + KT < T); }
    if (IDLE && dB == 0) { pfB += VS * CS; reload(accB, ltB, pfB + RS, s + 1 + KT < T); }
    const bool actA = !(IDLE && dA == 0), actB = !(IDLE && dB == 0);
    const bool ownA = (dA == 1) && (s + 1 < T), ownB = (dB == 1) && (s + 1 < T);
    double* const pkn = &S.PK[(s + 1) % LU_R][0][0];
    if (ownA || ownB) { wait_slot_free(s + 1); stage_edge(s + 1, (ownA ? pfA : pfB) + RS); }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const bool fvalid = (s + KT < T);
    if (s + 2 + KT < T) {   // L2 prefetch, two steps ahead
      if (dA >= 2) prefetch_tiles(pfA + 2 * RS, 0, 1);
      if (dB >= 2) prefetch_tiles(pfB + 2 * RS, 0, 1);
      if (!IDLE && dA == 2) prefetch_tiles(pfA + (2 - KT) * RS + KT * CS, RS, KT + 1);
      if (!IDLE && dB == 2) prefetch_tiles(pfB + (2 - KT) * RS + KT * CS, RS, KT + 1);
    }
    mbar_wait(xfull_bar(s), par);       // D_s^-1 is in shared memory
    mbar_wait(tfull_bar(s, 0), par);    // package tiles are consumed as their owner publishes them
    const uint32_t pk = smem_u32(&S.PK[slot][0][0] + 2 * lane);
    double2 afn = lds_v2(pk);           // package tile 0 = -A~(s+1, s)
    // ---------------- Ub(s, c) = D_s^-1 A~(s, c) for both columns, held as C fragments of Ub^T ----------------
    double2 wA = make_double2(0.0, 0.0), wB = make_double2(0.0, 0.0);
    {
      const double2 xc = *reinterpret_cast<const double2*>(&S.XC[slot][2 * lane]);
      if (actA) dmma884(wA.x, wA.y, accA[0].x, xc.x);
      if (actB) dmma884(wB.x, wB.y, accB[0].x, xc.x);
      if (actA) dmma884(wA.x, wA.y, accA[0].y, xc.y);
      if (actB) dmma884(wB.x, wB.y, accB[0].y, xc.y);
      if (!REV && actA && s + dA < T) stT(pfA - KT * RS, wA);
      if (!REV && actB && s + dB < T) stT(pfB - KT * RS, wB);
    }
    // ---------------- trailing update of both columns: A~(s+i, c)^T -= Ub^T A~(s+i, s)^T ----------------
    const bool give = (s + 2 < T);
    // tiles of update(s) the lookahead warp needs for D_{s+2}: t1 = row s+1, t2 = row s+2 of the column at distance d
    auto hand_over = [&](int d, const double2& t1, const double2& t2) {
      if (d == 1) {
        stT_s(S.tP[s & 1], t2);
        __syncwarp();
        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.tiles[s & 1]));
      } else if (d == 2) {
        *reinterpret_cast<double2*>(&S.tUt[s & 1][2 * lane]) = t1;
        *reinterpret_cast<double2*>(&S.tAt[s & 1][2 * lane]) = t2;
        stT_s(S.tA[s & 1], t2);
        __syncwarp();
        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.tiles[s & 1]));
      }
    };
    double2 tlA[NSM], tlB[NSM];
    double2 fA = make_double2(0.0, 0.0), fB = make_double2(0.0, 0.0);
#define LU2_TILE(acc_, tl_, f_, i_) (*((i_) < NR ? &acc_[(i_) < NR ? (i_) : 0] : ((i_) < KT ? &tl_[((i_) >= NR && (i_) < KT) ? (i_) - NR : 0] : &f_)))
#define TA(i_) LU2_TILE(accA, tlA, fA, i_)
#define TB(i_) LU2_TILE(accB, tlB, fB, i_)
    auto fetch_tail = [&](int i) {   // bring tile i of both columns into registers
      if (i >= NR && i < KT) {
        const int j = (i - NR < 0) ? 0 : (i - NR < NSM ? i - NR : 0);
        if (actA) tlA[j] = *reinterpret_cast<const double2*>(ltA + j * 64);
        if (actB) tlB[j] = *reinterpret_cast<const double2*>(ltB + j * 64);
      }
      if (i == KT) {   // entering row tiles (s+KT, c): from global memory (L2: prefetched two steps ago), transposed
        if (actA && fvalid) fA = ldT(pfA);
        if (actB && fvalid) fB = ldT(pfB);
        if (ownA || ownB) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tfull_bar(s + 1, KT - 1));   // the raw band-edge tile of package(s+1) has landed
        }
      }
    };
    auto load_operand = [&](int i) -> double2 {   // -A~(s+i, s): package tile i-1; the last one is the raw band-edge copy
      mbar_wait(tfull_bar(s, i - 1), par);
      if (i < KT) return lds_v2(pk + (i - 1) * 512);
      double2 ae = lds_v2(smem_u32(&S.PK[slot][KT - 1][REV ? 62 - 2 * lane : 2 * lane]));
      if (REV) ae = make_double2(ae.y, ae.x);
      return neg2(ae);
    };
    // tile i of a column is final: publish (next pivot column), write factors, move tail rows to their next slot
    auto finish = [&](int i, bool own, bool act, const double2& t, double* lt, double* pf) {
      if (own && i >= 2) {
        stT_s(pkn + (i - 2) * 64, neg2(t));
        __syncwarp();
        if (lane == 0) mbar_arrive(tfull_bar(s + 1, i - 2));
        if (!REV && s + i < T) stT(pf - (KT - i) * RS, t);
      }
      if (act && i > NR) *reinterpret_cast<double2*>(lt + (i - NR - 1) * 64) = t;   // rows NR+1..KT slide into the tail slots
    };
    constexpr int FD = 3;   // the entering tiles are requested FD iterations before their first use
    fetch_tail(1); fetch_tail(2);
    double2 af = afn;
    if (actA) dmma884(TA(1).x, TA(1).y, wA.x, af.x);
    if (actB) dmma884(TB(1).x, TB(1).y, wB.x, af.x);
#pragma unroll
    for (int i = 1; i <= KT; ++i) {
      if (i + 2 < KT) fetch_tail(i + 2);
      if (i + FD == KT) fetch_tail(KT);
      if (i < KT) afn = load_operand(i + 1);
      if (actA) dmma884(TA(i).x, TA(i).y, wA.y, af.y);
      if (actB) dmma884(TB(i).x, TB(i).y, wB.y, af.y);
      if (i < KT) {
        if (actA) dmma884(TA(i + 1).x, TA(i + 1).y, wA.x, afn.x);
        if (actB) dmma884(TB(i + 1).x, TB(i + 1).y, wB.x, afn.x);
      }
      // tile i-1 is final by now (its last DMMA was issued an iteration ago)
      if (i == 3 && give) { hand_over(dA, TA(1), TA(2)); hand_over(dB, TB(1), TB(2)); }
      if (i >= 2) { finish(i - 1, ownA, actA, TA(i - 1), ltA, pfA); finish(i - 1, ownB, actB, TB(i - 1), ltB, pfB); }
      af = afn;
    }
    finish(KT, ownA, actA, fA, ltA, pfA);
    finish(KT, ownB, actB, fB, ltB, pfB);
    __syncwarp();
    if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.empty[slot]));   // this warp no longer reads the slot of step s
    // window slide (rows NR+1.. went to shared memory in finish())
    if (actA) {
#pragma unroll
      for (int i = 1; i < NR; ++i) accA[i - 1] = accA[i];
      accA[NR - 1] = tlA[0];
    }
    if (actB) {
#pragma unroll
      for (int i = 1; i < NR; ++i) accB[i - 1] = accB[i];
      accB[NR - 1] = tlB[0];
    }
#undef TA
#undef TB
#undef LU2_TILE
    dA = (dA == 0) ? KT : dA - 1;   // (dA == 0 here only for the idle slot of odd KT)
    dB = (dB == 0) ? KT : dB - 1;
    pfA += RS;
    pfB += RS;
    if (s + 1 == T - KT) {
      if (dA < KT) schur_col(accA, ltA, dA);
      if (dB < KT) schur_col(accB, ltB, dB);
    }
    if (!IDLE) {   // even KT: the retired slot takes the entering column of the next step at once
      if (dA == 0) { pfA += KT * CS; reload(accA, ltA, pfA, s + 1 + KT < T); dA = KT; }
      if (dB == 0) { pfB += KT * CS; reload(accB, ltB, pfB, s + 1 + KT < T); dB = KT; }
    }
  }
}

template <int KT, bool REV>
static int launch_lu2_kt(spk_ctx* c, int grid, int first_part) {
  LuArgs a;
  a.band = c->band; a.schur = REV ? c->St : c->Sb; a.pstart = c->d_pstart;
  a.boost_count = (long long*)c->d_boost; a.tpr = c->L.tpr; a.tipT = c->tipT; a.first_part = first_part;
  a.boost_thr = c->opts.boost_rel * c->anorm_max;
  a.trace = nullptr;
  const size_t smem = sizeof(Lu2Smem<KT>);
  SPK_CUDA(c, cudaFuncSetAttribute(k_band_lu2<KT, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_band_lu2<KT, REV><<<grid, (Lu2Cfg<KT>::NW + 1) * 32, smem, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// two-columns-per-warp kernel for kt in [8, 16]; returns -1 when kt is outside its range
int spk_launch_lu2(spk_ctx* c, bool rev, int grid, int first_part) {
  switch (c->L.kt) {
#define CASE(K_) case K_: return rev ? launch_lu2_kt<K_, true>(c, grid, first_part) : launch_lu2_kt<K_, false>(c, grid, first_part);
    CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default: return -1;
  }
}
