// lu.cu -- per-partition banded block-LU (no pivoting, diagonal boosting) on FP64 tensor cores.
//
// Replaces PCSetUp(inner) = PETSc sparse LU of the AIJ band (/root/reference/src/matbanded.c:178).
//
// Factorisation computed (8x8 tiles, tile row/col indices I,J):   A = Lb * Ub  with
//     Lb(I,J) = A~(I,J)  (J <= I, D_I = A~(I,I)),      Ub(I,I) = I,  Ub(I,J) = D_I^-1 A~(I,J)  (J > I)
// where A~ is the Schur-updated matrix.  The band keeps the A~ tiles below the diagonal, the scaled
// Ub tiles above it and the EXPLICIT INVERSE D_I^-1 in the diagonal tile slot (that is what both this
// kernel and the triangular sweeps need; a scalar LU of D_I never has to be applied).  Mathematically
// this is the same elimination as the scalar no-pivot LU (same Schur complements, same pivots, same
// boosting rule), grouped by 8 pivots.
//
// One CTA per SPIKE partition, KT+1 warps (KT = ceil(K/8)), no CTA-wide barrier in the main loop:
//   * warp w < KT owns tile COLUMN J == w (mod KT) of the sliding KT x KT-tile trailing window; its
//     KT tiles live in registers as DMMA m8n8k4 accumulator fragments (2 doubles/lane/tile), so the
//     whole K x K window is register resident and every band entry is read and written once.  Tiles
//     entering the window are loaded straight into the fragments (one coalesced 16 B load per lane
//     and tile, L2-prefetched two steps ahead).
//   * per 8-pivot step s a column warp scales ITS OWN pivot-row tile, Ub(s,J) = D_s^-1 A~(s,J)
//     (2 DMMAs), then updates its column A~(I,J) -= A~(I,s) Ub(s,J) (2 DMMAs per tile, issued skewed so
//     that two accumulation chains are in flight).  The left operands come from the step PACKAGE: the
//     pivot-column tiles A~(s+1.., s) (the scaled pivot row is negated instead, once per step), published to
//     shared memory by the warp that owns column s WHILE it updates them during step s-1 and announced in
//     groups of LU_PUBG tiles (one mbarrier per group): the consumers of package(s) trail their producer by
//     a few tiles instead of waiting for its whole column.  That owner's update is the critical path of every
//     step; it also writes the new pivot column's factor tiles as it goes.  Packages live in a 4-deep ring; an
//     `empty` mbarrier per slot tells the producer when all column warps have left it.
//   * operand layouts: a DMMA contracts over k, so the SAME permutation of k may be applied to both
//     operands.  With k = 2*(lane%4)+h for k-chunk h, the left operand of M1*M2 is the accumulator
//     ("C") fragment of M1 itself and the right operand is the C fragment of M2^T.  Tiles are row-major
//     = C-fragment order in global memory, shared memory and registers, so package tiles and
//     accumulators are used as they are; the one layout change per warp and step is a register
//     transposition (shuffles) of the pivot-row tile, which yields Ub^T = U^T X^T (right operand of
//     every update) and Ub = X U (the stored factor).
//   * the INVERTER warp (warp KT) owns the only truly sequential part of the factorisation, the pivot-block
//     inverses.  The owner of the next pivot column updates row s+1 of its column first -- that tile is
//     D_{s+1} -- and hands it over through shared memory at once; the inverter inverts it ON THE TENSOR CORES
//     while the column warps run the rest of update(s): Newton-Schulz X <- X + (I - X D) X carried for X and
//     X^T (every operand a C fragment in registers, no shuffles, residual squares per step, stops below fp64
//     round-off), started from the entry-wise Jacobi iterate (2I - Dg^-1 D) Dg^-1 for diagonally dominant blocks,
//     else from an FP32 Gauss-Jordan without pivoting on the FP32 pipe, else the exact FP64 Gauss-Jordan with
//     the boosting rule (tiny pivots).  D_{s+1}^-1 goes to shared memory (own mbarrier) and to the diagonal
//     slot of the band; it is normally there before the column warps finish update(s).
//   * REV=true runs the same elimination on the row/column-reversed matrix (= bottom-up elimination
//     of the partition's first tipT tile rows) without storing factors: it only yields the top
//     Schur block S_t needed for the W^(t) spike tip.
// Outputs (FWD): block factors in place, the bottom Schur block S_b of each partition (for V^(b)),
// boosted-pivot count.
#include "lu_dev.cuh"

template <int KT>
struct LuSmem {
  double PK[LU_R][KT][64];   // package of step s (slot s % LU_R): A~(s+1+i, s), i = 0..KT-1, row-major (the last one a raw copy)
  double XC[LU_R][64];       // D_s^-1, row-major
  double LT[KT][2][64];      // per column warp: optional shared-memory tail of its column (LU_NSM_WIDE rows)
  double tD[2][64];          // next pivot block D_{u+1} = A~(u+1, u+1), handed over early in update(u), buffer u & 1
  double tDt[2][64];         // its transpose
  unsigned long long xfull[LU_R];       // 1 arrival: the inverter warp has published D_s^-1
  unsigned long long tfull[LU_R][KT];   // 1 arrival each: group g of package tiles (LU_PUBG tiles) published; [KT-1]: the band-edge tile
  unsigned long long empty[LU_R];   // KT arrivals: every column warp is done with the slot
};

template <int KT, bool REV, bool TRACE>
__global__ void __launch_bounds__((KT + 1) * 32, (KT >= 14 ? 1 : (KT >= 10 ? 2 : (KT >= 8 ? 3 : 4)))) k_band_lu(const LuArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LuSmem<KT>& S = *reinterpret_cast<LuSmem<KT>*>(smem_raw);
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived
  // from it (column bookkeeping, running pointers, ring slots) in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool is_inverter = (warp == KT);
  const int g = lane >> 2, tq = lane & 3;
  const int part = blockIdx.x + (REV ? a.first_part : 0);
  const int64_t t0 = a.pstart[part];
  const int64_t plen = a.pstart[part + 1] - t0;
  const int T = REV ? (int)(plen < a.tipT ? plen : a.tipT) : (int)plen;
  // steps to run: the tip window only wants the Schur block that is complete once step T-KT-1 has finished
  const int Tend = REV ? (T > KT ? T - KT : 0) : T;
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
  // tile (I,J) of the partition's diagonal block, zero outside it
  auto ld_tile = [&](int I, int J) -> double2 {
    return (I < T && J < T) ? ld_pair(tptr(I, J)) : make_double2(0.0, 0.0);
  };
  auto xfull_bar = [&](int s) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.xfull[s % LU_R]); };
  auto tfull_bar = [&](int s, int j) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.tfull[s % LU_R][j]); };
  auto empty_bar = [&](int s) -> uint64_t* { return reinterpret_cast<uint64_t*>(&S.empty[s % LU_R]); };
  // slot of step s is free again once every column warp has finished update(s - LU_R)
  auto wait_slot_free = [&](int s) {
    if (s >= LU_R) mbar_wait(empty_bar(s), (uint32_t)((s / LU_R - 1) & 1));
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < LU_R; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(&S.xfull[i]), 1);
      for (int j = 0; j < KT; ++j) mbar_init(reinterpret_cast<uint64_t*>(&S.tfull[i][j]), 1);
      mbar_init(reinterpret_cast<uint64_t*>(&S.empty[i]), KT);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (is_inverter) {
    // =========================== inverter warp ============================================
    const double thr = a.boost_thr, rthr = 1.0 / a.boost_thr;
    int nboost = 0;
    double2 x;
    for (int s = 0; s < Tend; ++s) {
      LU_TR(8);
      double2 d, dt;
      if (s == 0) {
        d = ld_pair(tptr(0, 0));
        dt = cfrag_transpose(d, g, tq);
      } else {   // D_s: the first tile the owner of column s updated in step s-1, handed over at once
        named_bar_sync(LU_BAR_TILES + ((s - 1) & 1), 64);
        d = *reinterpret_cast<const double2*>(&S.tD[(s - 1) & 1][2 * lane]);
        dt = *reinterpret_cast<const double2*>(&S.tDt[(s - 1) & 1][2 * lane]);
      }
      LU_TRV(9, d.x);
      LU_TRV(10, d.x);
      // pivot-block inverse on the tensor cores: Newton-Schulz from (1) the entry-wise Jacobi iterate (diagonally
      // dominant blocks: no elimination at all), else from (2) an FP32 Gauss-Jordan without pivoting on the
      // FP32 pipe, else (3) the exact FP64 Gauss-Jordan with the boosting rule
      double2 xt;
      jacobi_start8(d, dt, x, xt, g, tq);
      LU_TRV(14, x.x);
      if (!ns_refine8(d, dt, x, xt, g, tq)) {
        const float2 xf = gj8_f32_cfrag(d, g, tq);
        const float2 xft = cfrag_transpose_f(xf, g, tq);
        x = make_double2(f2d_bits(xf.x), f2d_bits(xf.y));
        xt = make_double2(f2d_bits(xft.x), f2d_bits(xft.y));
        if (!ns_refine8(d, dt, x, xt, g, tq)) x = gj8_cfrag(d, g, tq, thr, rthr, nboost);
      }
      LU_TRV(13, x.x);
      wait_slot_free(s);
      *reinterpret_cast<double2*>(&S.XC[s % LU_R][2 * lane]) = x;
      LU_TRV(11, x.x);
      __syncwarp();
      if (lane == 0) mbar_arrive(xfull_bar(s));   // D_s^-1 is published
      if (!REV) *reinterpret_cast<double2*>(tptr(s, s) + a.dst_off + 2 * lane) = x;  // factor output
      LU_TR(12);
    }
    if (lane == 0 && nboost) atomicAdd((unsigned long long*)a.boost_count, (unsigned long long)nboost);
    return;
  }

  // =========================== column warps ================================================
  // accT[i] = C fragment (row-major) of A~(s+i, c), i = 0..KT-1, for the warp's current column c = s + cj
  // (cj = jrel, or KT for the warp whose column retired at this step and that has taken over the entering
  // column s+KT).  All tiles move between memory and registers as they are (one 16 B access per lane); the one
  // layout change per step is the register transposition of the pivot-row tile accT[0]: with U^T in hand,
  // Ub^T = U^T X^T (right operand of every update) and Ub = X U (factor output) are two DMMA pairs.
  constexpr int SGN = REV ? -1 : 1;
  const int RS = SGN * (tpr - 1) * SPK_TILE_ELEMS;   // one tile row down, same column (doubles)
  constexpr int CS = SGN * SPK_TILE_ELEMS;           // one tile column to the right
  const int l0 = 16 * tq + g;
  auto ldT = [&](const double* tile) -> double2 { return ld_pair(tile); };
  auto stT = [&](double* tile, const double2& v) { *reinterpret_cast<double2*>(tile + 2 * lane) = v; };   // (!REV only)
  auto stT_s = [&](double* tile, const double2& v) { *reinterpret_cast<double2*>(tile + 2 * lane) = v; };
  auto stTr_s = [&](double* tile, const double2& v) { tile[l0] = v.x; tile[l0 + 8] = v.y; };   // shared-memory tile <- M^T

  // rows 0..NR-1 of the column live in registers, the NSM newest rows in the warp's shared-memory slots
  // (C-fragment order: every lane reads back its own 16 B).  KT = 13 would need 52 accumulator registers;
  // with the 72 that two CTAs per SM leave, the compiler then spills accumulators and cannot keep a
  // package tile in flight ahead of the tensor pipe.
  constexpr int NSM = (KT >= 12) ? LU_NSM_WIDE : 0;   // 0: the whole column in registers
  constexpr int NR = KT - NSM;
  double2 accT[NR];
  double* const lt = &S.LT[warp][0][2 * lane];   // slot j at lt + j*64
  auto col_tile = [&](int i, int J) -> double2 { return (i < T && J < T) ? ldT(tptr(i, J)) : make_double2(0.0, 0.0); };
#pragma unroll
  for (int i = 0; i < NR; ++i) accT[i] = col_tile(i, warp);
#pragma unroll
  for (int j = 0; j < NSM; ++j) *reinterpret_cast<double2*>(lt + j * 64) = col_tile(NR + j, warp);
  int jrel = warp;                   // (column owned) - s, taken mod KT
  double* pf = tptr(KT, warp);       // running pointer: tile (s+KT, c)

  // Package of pivot column sn: tiles 0..KT-2 = A~(sn+1+j, sn) (logical row-major), tile KT-1 = the band-edge
  // tile (sn+KT, sn), which no update ever touched: a RAW copy (stored orientation, not negated) made by cp.async.
  auto stage_edge = [&](int sn, const double* src) {
    double* dst = &S.PK[sn % LU_R][KT - 1][2 * lane];
    if (sn + KT < T) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + 2 * lane) : "memory");
    else *reinterpret_cast<double2*>(dst) = make_double2(0.0, 0.0);
  };
  // the warp whose column retired recycles its registers for the entering column sn+KT (pf = tile (sn+KT, sn))
  auto reload_entering = [&](int sn) {
    const bool cv = (sn + KT < T);
    pf += KT * CS;                   // tile (sn+KT, sn+KT)
#pragma unroll
    for (int i = 0; i < NR; ++i) accT[i] = cv ? ldT(pf - (KT - i) * RS) : make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < NSM; ++j) *reinterpret_cast<double2*>(lt + j * 64) = cv ? ldT(pf - (NSM - j) * RS) : make_double2(0.0, 0.0);
  };
  // L2 prefetch of cnt tiles starting at p, stride `step` doubles (4 lines of 128 B per tile)
  auto prefetch_tiles = [&](const double* p, int step, int cnt) {
    for (int l = lane; l < 4 * cnt; l += 32) {
      const double* q = p + (l >> 2) * step + (l & 3) * 16;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    }
  };

  // when step T-KT is about to start the warps hold the trailing Schur complement of the partition
  // (rows/cols T-KT..T-1); called before the pivot column's owner recycles its registers
  const int kp = KT * 8;
  auto schur_out = [&]() {
    double* out = a.schur + (int64_t)part * kp * kp;
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const double2 t = (i < NR) ? accT[i < NR ? i : 0] : *reinterpret_cast<const double2*>(lt + (i < NR ? 0 : i - NR) * 64);
      const int r = 8 * i + g, cc = 8 * jrel + 2 * tq;
      if (!REV) {
        *reinterpret_cast<double2*>(out + (int64_t)r * kp + cc) = t;
      } else {
        out[(int64_t)(kp - 1 - r) * kp + (kp - 1 - cc)] = t.x;
        out[(int64_t)(kp - 1 - r) * kp + (kp - 2 - cc)] = t.y;
      }
    }
  };
  if (T == KT) schur_out();
  if (jrel == 0) {   // nothing is eliminated yet: column 0 is published as loaded (its factor tiles are already in place)
    stage_edge(0, pf);
    asm volatile("cp.async.commit_group;" ::: "memory");
    double* pk0 = &S.PK[0][0][0];
#pragma unroll
    for (int i = 1; i < KT; ++i) {
      const double2 t = (i < NR) ? accT[i < NR ? i : 0] : *reinterpret_cast<const double2*>(lt + (i < NR ? 0 : i - NR) * 64);
      stT_s(pk0 + (i - 1) * 64, t);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (lane < KT) mbar_arrive(tfull_bar(0, lane));   // (all group barriers and the band-edge one; unused indices are harmless)
    reload_entering(0);
  }

  for (int s = 0; s < Tend; ++s) {
    const int cj = (jrel == 0) ? KT : jrel;
    const int slot = s % LU_R;
    // the warp that owns the NEXT pivot column publishes package(s+1) tile by tile while it updates them
    const bool own_next = (jrel == 1) && (s + 1 < Tend);
    double* const pkn = &S.PK[(s + 1) % LU_R][0][0];
    if (own_next) { wait_slot_free(s + 1); stage_edge(s + 1, pf + RS); }
    // ---- independent of the package: the entering row tile lands in shared memory; L2 prefetch two steps ahead
    const bool fvalid = (s + KT < T);   // (all columns of the window are inside the partition then)
    asm volatile("cp.async.commit_group;" ::: "memory");   // (the band-edge copy of the next package, if any)
    if (s + 2 + KT < T) {
      if (cj >= 2) prefetch_tiles(pf + 2 * RS, 0, 1);
      if (jrel == (KT > 2 ? 2 : KT - 1)) prefetch_tiles(pf + (2 - KT) * RS + (2 + KT - (KT > 2 ? 2 : KT - 1)) * CS, RS, KT + 1);
    }
    const double2 ut = cfrag_transpose(accT[0], g, tq);   // pivot-row tile, transposed in registers
    if (warp == 0) LU_TR(0);
    const uint32_t par = (uint32_t)((s / LU_R) & 1);
    mbar_wait(xfull_bar(s), par);       // D_s^-1 is in shared memory
    mbar_wait(tfull_bar(s, 0), par);    // package tiles are consumed as their owner publishes them
    if (warp == 0) LU_TR(1);
    // ---------------- Ub(s, c) = D_s^-1 A~(s, c), held as the C fragment of Ub^T = U^T X^T ----------------
    // package tiles come through volatile shared-memory loads, one tile ahead of the tensor pipe (the compiler
    // must not hoist a batch of them: the accumulators need the registers)
    const uint32_t pk = smem_u32(&S.PK[slot][0][0] + 2 * lane);
    double2 afn = lds_v2(pk);   // package tile 0 = A~(s+1, s)
    double2 w = make_double2(0.0, 0.0);
    {
      const double2 xc = *reinterpret_cast<const double2*>(&S.XC[slot][2 * lane]);
      dmma_cc(w, ut, xc);                        // Ub^T = U^T X^T
      if (!REV && s + cj < T) stT(pf + a.dst_off - KT * RS, cfrag_transpose(w, g, tq));   // Ub, the stored factor
      w = neg2(w);                               // the update subtracts: A~ += A~(.,s) (-Ub); the package holds +A~(.,s)
    }
    if (warp == 0) LU_TR(2);
    // ---------------- trailing update of the column: A~(s+i, c) -= A~(s+i, s) Ub(s, c) ----------------
    // the owner of the next pivot column hands the next pivot block D_{s+1} (row s+1 of its column, the first
    // tile it updates) to the inverting warp the moment it is final
    auto give_d = [&](const double2& t1) {
      stT_s(S.tD[s & 1], t1);
      stTr_s(S.tDt[s & 1], t1);
      named_bar_arrive(LU_BAR_TILES + (s & 1), 64);
    };
    // row s+i of column s+1 goes into package(s+1) as tile i-2 (row s+1 is the next pivot block itself);
    auto pub = [&](int i, const double2& t) {
      if (i >= 2 && own_next) {
        stT_s(pkn + (i - 2) * 64, t);
        if (!REV && s + i < T) stT(pf + a.dst_off - (KT - i) * RS, t);   // factor output of the new pivot column
        // tiles are announced in groups of LU_PUBG (one mbarrier per group: fewer warp syncs + arrivals on the
        // producer's in-order stream, which everybody else's next step hangs on)
        if ((i - 2) % LU_PUBG == LU_PUBG - 1 || i - 2 == KT - 2) {
          __syncwarp();
          if (lane == 0) mbar_arrive(tfull_bar(s + 1, (i - 2) / LU_PUBG));
        }
      }
    };
    // One skewed loop over the 13 tiles of the column (rows s+1 .. s+KT): the second k-chunk of tile i and the
    // first k-chunk of tile i+1 are issued back to back, so two accumulation chains are in flight per warp and
    // the operand load of tile i+1 overlaps the tensor-pipe latency of tile i.  Rows NR.. come from the
    // shared-memory tail, row s+KT is the entering tile; everything is resolved at compile time.
    double2 tl[NSM > 0 ? NSM : 1];
    double2 fT = make_double2(0.0, 0.0);
#define LU_TILE(i_) (*((i_) < NR ? &accT[(i_) < NR ? (i_) : 0] : ((i_) < KT ? &tl[((i_) >= NR && (i_) < KT) ? (i_) - NR : 0] : &fT)))
    auto fetch_tail = [&](int i) {   // make tile i available in registers (called two iterations before its first use)
      if (i >= NR && i < KT) tl[i - NR < 0 ? 0 : (i - NR < NSM ? i - NR : 0)] = *reinterpret_cast<const double2*>(lt + (i - NR) * 64);
      if (i == KT) {
        // entering row tile (s+KT, c): straight from global memory (L2: prefetched two steps ago), transposed
        fT = fvalid ? ldT(pf) : make_double2(0.0, 0.0);
        if (own_next) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tfull_bar(s + 1, KT - 1));   // the raw band-edge tile of package(s+1) has landed
        }
      }
    };
    auto load_operand = [&](int i) -> double2 {   // A~(s+i, s): package tile i-1; the last one is the raw band-edge copy
      if (i == KT) mbar_wait(tfull_bar(s, KT - 1), par);                                  // the raw band-edge tile
      else if ((i - 1) % LU_PUBG == 0) mbar_wait(tfull_bar(s, (i - 1) / LU_PUBG), par);   // first tile of a group
      if (i < KT) return lds_v2(pk + (i - 1) * 512);
      double2 ae = lds_v2(smem_u32(&S.PK[slot][KT - 1][REV ? 62 - 2 * lane : 2 * lane]));
      if (REV) ae = make_double2(ae.y, ae.x);
      return ae;
    };
    constexpr int FD = (KT > 6) ? 6 : 2;   // the entering tile is requested FD iterations before its first use
    fetch_tail(1); fetch_tail(2);
    if (KT <= FD) fetch_tail(KT > 2 ? KT : 0);
    double2 af = afn;
    // Two operand buffers, af / afn.  The operand of tile i+2 is requested right after the last DMMA that reads the
    // registers it lands in (the second k-chunk of tile i), so TWO DMMAs sit between a shared-memory load and its first
    // use; requesting it an iteration later (one DMMA in between) left the tensor pipe waiting on the short scoreboard
    // (ncu source view: 28 % of the samples on the DMMA behind each LDS; 8.83 -> 8.59 ms at C3).
    if (KT >= 2) afn = load_operand(2);
    dmma884(LU_TILE(1).x, LU_TILE(1).y, af.x, w.x);
#pragma unroll
    for (int i = 1; i <= KT; ++i) {
      if (i + 2 < KT) fetch_tail(i + 2);
      if (KT > FD && i + FD == KT) fetch_tail(KT);
      dmma884(LU_TILE(i).x, LU_TILE(i).y, af.y, w.y);
      if (i + 2 <= KT) af = load_operand(i + 2);
      if (i < KT) dmma884(LU_TILE(i + 1).x, LU_TILE(i + 1).y, afn.x, w.x);
      if (i == 1 && own_next) give_d(LU_TILE(1));
      if (i >= 2) pub(i - 1, LU_TILE(i - 1));
      const double2 nx = af; af = afn; afn = nx;
    }
    pub(KT, fT);
#undef LU_TILE
    __syncwarp();
    if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.empty[slot]));   // this warp no longer reads the slot of step s
    // window slide
#pragma unroll
    for (int i = 1; i < NR; ++i) accT[i - 1] = accT[i];
    accT[NR - 1] = (NSM > 0) ? tl[0] : fT;
#pragma unroll
    for (int j = 1; j < NSM; ++j) *reinterpret_cast<double2*>(lt + (j - 1) * 64) = tl[j];
    if (NSM > 0) *reinterpret_cast<double2*>(lt + (NSM > 0 ? NSM - 1 : 0) * 64) = fT;
    jrel = (jrel == 0) ? KT - 1 : jrel - 1;
    pf += RS;
    if (warp == 0) LU_TR(4);
    if (s + 1 == T - KT) schur_out();
    if (own_next) reload_entering(s + 1);
    if (warp == 0) LU_TR(5);
  }
}

// --------------------------------------------------------------------------------------------
template <int KT, bool REV, bool TRACE>
static int launch_lu_kt(spk_ctx* c, int grid, int first_part) {
  LuArgs a;
  // out of place when the unfactored band is kept (and not equilibrated in place afterwards): read the original,
  // write the factors -- the same 2B of traffic, and the factorisation can be repeated without restoring anything
  double* src = spk_lu_source(c);
  a.band = src; a.dst_off = (long long)(c->band - src);
  a.schur = REV ? c->St : c->Sb; a.pstart = c->d_pstart;
  a.boost_count = (long long*)c->d_boost; a.tpr = c->L.tpr; a.tipT = c->tipT; a.first_part = first_part;
  a.boost_thr = c->opts.boost_rel * c->anorm_max;
  a.trace = (long long*)c->lu_trace;
  const size_t smem = sizeof(LuSmem<KT>);
  SPK_CUDA(c, cudaFuncSetAttribute(k_band_lu<KT, REV, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_band_lu<KT, REV, TRACE><<<grid, (KT + 1) * 32, smem, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

template <bool REV>
static int launch_lu(spk_ctx* c, int grid, int first_part) {
  if (!REV && c->lu_trace && c->L.kt == 13) return launch_lu_kt<13, false, true>(c, grid, first_part);  // tools/lu_trace.py
  switch (c->L.kt) {
#define CASE(K_) case K_: return launch_lu_kt<K_, REV, false>(c, grid, first_part);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default:
      SPK_SET_ERR(c, "half-bandwidth %d (kt=%d) outside the supported range 1..%d", c->L.k, c->L.kt, 8 * SPK_MAX_KT);
      return SPK_ERR_UNSUPPORTED;
  }
}

int spk_launch_lu(spk_ctx* c) { return c->wide ? spk_wide_band_lu(c) : launch_lu<false>(c, c->P, 0); }

// bottom-up windows for W^(t): partitions 1..P-1 (and partition 0 when a left-neighbour rank exists)
int spk_launch_ul_tips(spk_ctx* c) {
  if (c->wide) return spk_wide_ul_windows(c);
  const int first = (c->opts.rank > 0) ? 0 : 1;
  const int grid = c->P - first;
  if (grid <= 0) return SPK_OK;
  return launch_lu<true>(c, grid, first);
}
