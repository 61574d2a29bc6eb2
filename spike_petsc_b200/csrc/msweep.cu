// msweep.cu -- partition sweeps for SEVERAL right-hand sides at once (spk_solve with nrhs >= 2; replaces looping
// PCApply / MatSolve over the columns, /root/reference/src/matbanded.c:190, and is what MatMatSolve would bind to).
//
// g_i = A_i^-1 B_i for a block of right-hand sides reads every factor tile ONCE and turns the per-row work into
// 8x8 tile products on the FP64 tensor cores: with the right-hand sides taken 8 columns at a time, the unknowns
// of tile row I form an 8x8 block Y_I and
//     forward :  Y_I = D_I^-1 ( B_I - sum_{d=1..KT} Lb(I,I-d) Y_{I-d} )          (explicit D^-1 in the diagonal slot)
//     backward:  X_I =          Y_I - sum_{d=1..KT} Ub(I,I+d) X_{I+d}            (unit block diagonal)
// are DMMA m8n8k4 pairs whose left operands are the factor tiles exactly as stored (row-major = accumulator
// fragment order, see lu.cu) and whose right operands are the transposed blocks solved before, which never leave
// the registers of the warp that owns the column group.
//
// One CTA per partition: up to MS_GROUPS warps, each owning 8 right-hand-side columns and running the whole
// recurrence on its own (no barrier between the warps), plus one producer thread that streams the tile rows through
// a ring of cp.async.bulk stages (full/empty mbarriers).  Per tile row and warp: 2*KT + 2 DMMAs in four independent
// accumulation chains, two register transpositions, two 8-byte loads and stores per lane for the vectors
// (prefetched MS_PF rows ahead).  The band is read once per 8*MS_GROUPS columns instead of once per column.
#include "lu_dev.cuh"

#define MS_NST 6          // stage ring depth
#define MS_GROUPS 4       // column groups (warps) per CTA: 32 right-hand sides per pass over the band
#define MS_PF 4           // vector rows prefetched ahead of the recurrence

enum { MSWEEP_MAIN = 0, MSWEEP_CORR = 1 };
struct MSweepArgs {
  const double* band; int tpr;
  const int64_t* pstart;
  const double* in;     // MAIN: right-hand sides, column r at in + r*ld
  double* x;            // MAIN: solutions (may alias `in`); CORR: vectors being corrected
  int64_t ld, n;
  int nrhs, col0;       // this launch handles columns col0 .. col0 + 8*MS_GROUPS - 1 (those < nrhs)
  // window corrections (solve.cu, k_sweep CORR, for 8 columns per warp): x_i -= A_i^-1 [r_top; 0] + A_i^-1 [0; r_bot]
  // restricted to the truncation window; CTA = 2*partition + side
  int mode, P, tipT;
  const double* tips;   // column r: r_top at tips + r*tip_stride (P*kp), r_bot right behind it
  size_t tip_stride;
  double* work;         // forward results of the window sweeps, column r at work + r*ld_work
  int64_t ld_work;
  int has_left, has_right;   // neighbour ranks present: partition 0's top / partition P-1's bottom correction is active
};

template <int KT>
struct MSweepSmem {
  double stage[MS_NST][KT + 1][64];
  unsigned long long full[MS_NST];
  unsigned long long empty[MS_NST];
};

template <int KT, int MODE>
__global__ void __launch_bounds__((MS_GROUPS + 1) * 32) k_msweep(const MSweepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MSweepSmem<KT>& S = *reinterpret_cast<MSweepSmem<KT>*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int live = a.nrhs - a.col0;
  const int ngroups = live >= 8 * MS_GROUPS ? MS_GROUPS : (live + 7) / 8;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MS_NST; ++i) {
      mbar_init(reinterpret_cast<uint64_t*>(&S.full[i]), 1);
      mbar_init(reinterpret_cast<uint64_t*>(&S.empty[i]), ngroups);
    }
    fence_mbar_init();
  }
  __syncthreads();
  constexpr bool corr = (MODE == MSWEEP_CORR);
  const int p = corr ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int64_t t0 = a.pstart[p], t1 = a.pstart[p + 1];
  // forward sweep over tile rows [f0, f1) ascending, then backward over [b0, b1) descending
  int64_t f0 = t0, f1 = t1, b0 = t0, b1 = t1;
  bool use_top = false, use_bot = false;
  if (corr) {   // same jobs as k_sweep's correction mode (single rank: no remote neighbours)
    const int side = blockIdx.x & 1;
    const int64_t plen = t1 - t0;
    const bool top_on = (p > 0) || a.has_left, bot_on = (p < a.P - 1) || a.has_right;   // neighbour ranks: sharded contexts
    const bool full = 2 * (int64_t)a.tipT > plen;
    const int64_t W = full ? plen : a.tipT;
    if (full) {
      if (side == 1 || (!top_on && !bot_on)) return;
      use_top = top_on; use_bot = bot_on; b0 = t0; b1 = t1; f0 = top_on ? t0 : t1 - KT; f1 = t1;
    } else if (side == 0) {
      if (!top_on) return;
      use_top = true; b0 = t0; b1 = t0 + W; f0 = t0; f1 = b1;
    } else {
      if (!bot_on) return;
      use_bot = true; b0 = t1 - W; b1 = t1; f0 = t1 - KT; f1 = t1;
    }
  }
  const int nf = (int)(f1 - f0), nb = (int)(b1 - b0);
  // iteration gi = 0 .. nf+nb-1: forward rows f0.., then backward rows b1-1..; stage gi % MS_NST, use gi / MS_NST
  if (warp == MS_GROUPS) {
    if (lane == 0) {
      for (int gi = 0; gi < nf + nb; ++gi) {
        const int st = gi % MS_NST, use = gi / MS_NST;
        if (use >= 1) mbar_wait(reinterpret_cast<uint64_t*>(&S.empty[st]), (uint32_t)((use - 1) & 1));
        const bool fwd = gi < nf;
        const int64_t I = fwd ? f0 + gi : b1 - 1 - (gi - nf);
        const uint32_t bytes = (uint32_t)((fwd ? KT + 1 : KT) * 512);
        uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[st]);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(&S.stage[st][0][0], a.band + (I * a.tpr + (fwd ? 0 : KT + 1)) * SPK_TILE_ELEMS, bytes, bar);
      }
    }
    return;
  }
  if (warp >= ngroups) return;

  const int cA = a.col0 + 8 * warp + 2 * tq;
  const bool vA = cA < a.nrhs, vB = cA + 1 < a.nrhs;
  const size_t oA = (size_t)cA * (size_t)a.ld, oB = oA + (size_t)a.ld;
  const size_t wA = (size_t)cA * (size_t)a.ld_work, wB = wA + (size_t)a.ld_work;
  const int kp = KT * 8;
  const double* rtA = corr ? a.tips + (size_t)cA * a.tip_stride + (size_t)p * kp : nullptr;   // r_top of this partition, column cA
  const double* rtB = corr ? rtA + a.tip_stride : nullptr;
  const size_t rb_off = (size_t)a.P * kp;                                         // r_bot sits P*kp behind r_top
  // right-hand side block of tile row I for the forward sweep / input block for the backward sweep
  auto ldv = [&](bool fwd, int64_t I) -> double2 {
    const int64_t row = I * 8 + g;
    double2 r = make_double2(0.0, 0.0);
    if (!corr) {
      const double* v = fwd ? a.in : a.x;
      if (row < a.n) { if (vA) r.x = v[oA + row]; if (vB) r.y = v[oB + row]; }
    } else if (fwd) {   // [r_top; 0; r_bot]: only the first / last KT tile rows of the partition carry a right-hand side
      if (use_top && I < t0 + KT) { const int64_t e = row - t0 * 8; if (vA) r.x += rtA[e]; if (vB) r.y += rtB[e]; }
      if (use_bot && I >= t1 - KT) { const int64_t e = row - (t1 - KT) * 8; if (vA) r.x += rtA[rb_off + e]; if (vB) r.y += rtB[rb_off + e]; }
    } else if (I >= f0) {   // forward results (rows above the forward range had a zero right-hand side)
      if (vA) r.x = a.work[wA + row]; if (vB) r.y = a.work[wB + row];
    }
    return r;
  };
  auto stv = [&](bool fwd, int64_t I, const double2& y) {
    const int64_t row = I * 8 + g;
    if (!corr) {
      if (row < a.n) { if (vA) a.x[oA + row] = y.x; if (vB) a.x[oB + row] = y.y; }
    } else if (fwd) {
      if (vA) a.work[wA + row] = y.x; if (vB) a.work[wB + row] = y.y;
    } else if (row < a.n) {   // the backward result is the correction itself
      if (vA) a.x[oA + row] -= y.x; if (vB) a.x[oB + row] -= y.y;
    }
  };
  double2 prev[KT];   // prev[d-1] = C fragment of (-Y_{I-d})^T : right operand of the tile product with distance d
  double2 rh[MS_PF];
  int gi = 0;
#pragma unroll 1
  for (int dir = 0; dir < 2; ++dir) {
    const bool fwd = (dir == 0);
    const int nrows = fwd ? nf : nb;
    const int64_t Ifirst = fwd ? f0 : b1 - 1;
    const int64_t step = fwd ? 1 : -1;
    // only blocks solved earlier in the same sweep contribute (tiles reaching outside the partition hold the
    // coupling blocks; they meet zero blocks)
#pragma unroll
    for (int d = 0; d < KT; ++d) prev[d] = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < MS_PF; ++j) rh[j] = (j < nrows) ? ldv(fwd, Ifirst + step * j) : make_double2(0.0, 0.0);
#pragma unroll 1
    for (int it = 0; it < nrows; ++it, ++gi) {
      const int64_t I = Ifirst + step * it;
      double2 acc0 = rh[0], acc1 = make_double2(0.0, 0.0), acc2 = acc1, acc3 = acc1;
#pragma unroll
      for (int j = 0; j + 1 < MS_PF; ++j) rh[j] = rh[j + 1];
      rh[MS_PF - 1] = (it + MS_PF < nrows) ? ldv(fwd, I + step * MS_PF) : make_double2(0.0, 0.0);
      const int st = gi % MS_NST;
      mbar_wait(reinterpret_cast<uint64_t*>(&S.full[st]), (uint32_t)((gi / MS_NST) & 1));
      const uint32_t base = smem_u32(&S.stage[st][0][2 * lane]);
      // forward: stage tile t holds distance d = KT - t (t = KT: D^-1); backward: stage tile t holds d = t + 1.
      // Farthest first, the product with the block solved one row ago last.
      double2 tl = lds_v2(base + (fwd ? 0 : (KT - 1) * 512));
#pragma unroll
      for (int d = KT; d >= 1; --d) {
        const double2 cur = tl;
        if (d > 1) tl = lds_v2(base + (fwd ? (KT - d + 1) : (d - 2)) * 512);
        double2& acc = ((KT - d) & 3) == 0 ? acc0 : ((KT - d) & 3) == 1 ? acc1 : ((KT - d) & 3) == 2 ? acc2 : acc3;
        dmma_cc(acc, cur, prev[d - 1]);
      }
      double2 dinv = make_double2(0.0, 0.0);
      if (fwd) dinv = lds_v2(base + KT * 512);
      __syncwarp();
      if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.empty[st]));
      double2 y = make_double2((acc0.x + acc1.x) + (acc2.x + acc3.x), (acc0.y + acc1.y) + (acc2.y + acc3.y));
      if (fwd) {
        const double2 accT = cfrag_transpose(y, g, tq);
        y = make_double2(0.0, 0.0);
        dmma_cc(y, dinv, accT);                    // Y_I = D_I^-1 (...)
      }
      stv(fwd, I, y);
#pragma unroll
      for (int d = KT - 1; d >= 1; --d) prev[d] = prev[d - 1];
      prev[0] = cfrag_transpose(neg2(y), g, tq);
    }
  }
}

template <int KT>
static int launch_msweep_kt(spk_ctx* c, const MSweepArgs& a) {
  const size_t smem = sizeof(MSweepSmem<KT>);
  if (a.mode == MSWEEP_CORR) {
    SPK_CUDA(c, cudaFuncSetAttribute(k_msweep<KT, MSWEEP_CORR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_msweep<KT, MSWEEP_CORR><<<2 * c->P, (MS_GROUPS + 1) * 32, smem, c->stream>>>(a);
  } else {
    SPK_CUDA(c, cudaFuncSetAttribute(k_msweep<KT, MSWEEP_MAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_msweep<KT, MSWEEP_MAIN><<<c->P, (MS_GROUPS + 1) * 32, smem, c->stream>>>(a);
  }
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

static int launch_msweep_cols(spk_ctx* c, MSweepArgs a, int nrhs) {
  for (int col0 = 0; col0 < nrhs; col0 += 8 * MS_GROUPS) {
    a.nrhs = nrhs; a.col0 = col0;
    int rc;
    switch (c->L.kt) {
#define CASE(K_) case K_: rc = launch_msweep_kt<K_>(c, a); break;
      CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
      default: SPK_SET_ERR(c, "unsupported kt=%d", c->L.kt); return SPK_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
  }
  return SPK_OK;
}

// g = D^-1 b for nrhs columns (column r at b + r*ld / x + r*ld), 8*MS_GROUPS columns per pass over the band
int spk_launch_msweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld) {
  if (c->wide) return spk_wide_main_sweep(c, b, x, nrhs, ld);
  MSweepArgs a{};
  a.band = c->band; a.tpr = c->L.tpr; a.pstart = c->d_pstart; a.in = b; a.x = x; a.ld = ld; a.n = c->L.n;
  a.mode = MSWEEP_MAIN; a.P = c->P; a.tipT = c->tipT;
  return launch_msweep_cols(c, a, nrhs);
}

// window corrections of all columns: tips = coupling right-hand sides written by spk_launch_reduced_solve_multi,
// work = nrhs columns of padded length ld_work (forward results of the window sweeps)
int spk_launch_mcorrections(spk_ctx* c, double* x, int nrhs, int64_t ld, const double* tips, double* work, int64_t ld_work) {
  if (c->wide) return spk_wide_corrections(c, x, nrhs, ld, tips, tips + (size_t)c->P * c->kp, 2 * (size_t)c->P * c->kp, work, ld_work);
  MSweepArgs a{};
  a.band = c->band; a.tpr = c->L.tpr; a.pstart = c->d_pstart; a.in = nullptr; a.x = x; a.ld = ld; a.n = c->L.n;
  a.mode = MSWEEP_CORR; a.P = c->P; a.tipT = c->tipT;
  a.tips = tips; a.tip_stride = 2 * (size_t)c->P * c->kp; a.work = work; a.ld_work = ld_work;
  a.has_left = c->opts.rank > 0; a.has_right = c->opts.rank + 1 < c->opts.nranks;
  return launch_msweep_cols(c, a, nrhs);
}
