// solve.cu -- the SPIKE preconditioner apply  x = B^{-1} b   (replaces PCApply(inner) = PETSc
// MatSolve_SeqAIJ, /root/reference/src/matbanded.c:190).
//
//   (1) k_sweep<MAIN>   g_i = A_i^{-1} b_i : forward (L) and backward (U) sweeps of every partition,
//                       factors streamed exactly once through a cp.async.bulk ring.
//   (2) k_reduced_solve per interface: (I - W V) x_t = g_t - W g_b ; x_b = g_b - V x_t ; coupling
//                       right-hand sides r_top = C x_b, r_bot = B x_t for the corrections.
//   (3) k_sweep<CORR>   x_i = g_i - A_i^{-1}[r_top;0] - A_i^{-1}[0;r_bot], restricted to the
//                       truncation window (tipT tile rows) where the spikes have not yet decayed;
//                       tipT = whole partition reproduces the classical second full sweep.
//
// Sweep kernel anatomy (one CTA per partition / correction job, 5 warps):
//   warp 0 ("near")  carries the sequential recurrence  y_I = D_I^-1 (c_I - Lb(I,I-1) y_{I-1})
//                    with 8x8 tiles spread over the lanes and warp-shuffle reductions;
//   warps 1-3 ("far") accumulate the part of the dot products that only needs y_{<=I-2} for the
//                    NEXT tile row (warp-shuffle reductions over the 4 lanes of a row);
//   warp 4 keeps SW_NST_MAIN / SW_NST_CORR tile rows in flight with bulk async copies (mbarrier tx-count).
#include "common.cuh"
#include <algorithm>

#define SW_NST_MAIN 8         // stage ring depth of the partition sweeps (2 CTAs/SM resident)
#define SW_NST_CORR 6         // ... of the window corrections: 46 KB of shared memory, 4 CTAs/SM = all 2P jobs resident
#define SW_THREADS 160          // near warp, three far warps, one copy warp
#define SW_COPY_THREAD 128

enum { SWEEP_MAIN = 0, SWEEP_CORR = 1 };

// -DSW_TRACE (tools/sweep_trace.py, tools/build_variant.sh): clock64 stamps of CTA 0, forward sweep iterations
// 32..63, per role (near warp: 8 points, far warp 1: 4 points, copy warp: 3 points).  Not in the product build.
#ifdef SW_TRACE
#ifndef SW_TRACE_CTA
#define SW_TRACE_CTA 0   // 0: partition sweep of partition 0; 2: correction job (partition 1, top window) -- the last launch wins
#endif
__device__ long long g_sw_trace[32][16];
#define SW_TR(slot_) do { if (tr_on && (threadIdx.x & 31) == 0) g_sw_trace[it - 32][slot_] = clock64(); } while (0)
extern "C" int spk_debug_sweep_trace(long long* out) { return cudaMemcpyFromSymbol(out, g_sw_trace, sizeof(g_sw_trace)) == cudaSuccess ? 0 : 1; }
#else
#define SW_TR(slot_) do { } while (0)
#endif

struct SweepArgs {
  const double* band; int tpr;
  const int64_t* pstart; int P;
  int mode;
  const double* in;    // MAIN: right-hand side b
  double* x;           // MAIN: output (y then x, in place); CORR: vector being corrected
  double* work;        // CORR: scratch vector (n_padded) holding the correction w
  const double* rtop;  // CORR: P * kp   r_top of partition p   (C_p x_b(p-1))
  const double* rbot;  // CORR: P * kp   r_bot of partition p   (B_p x_t(p+1))
  int tipT;            // CORR window (tile rows)
  int has_left, has_right;  // remote neighbours present (multi-GPU): partition 0 top / P-1 bottom active
  int64_t n;           // rows of the user vectors (padded rows are neither read nor written)
};

template <int KT, int SW_NST>
struct SweepSmem {
  // forward : [0..KT-1] = Lb tiles d=-KT..-1, [KT] = D^-1 (diagonal slot), [KT+1] = right-hand-side block (8 doubles)
  // backward: [0..KT-1] = Ub tiles d=1..KT (unit block diagonal),         [KT+1] = right-hand-side block
  double stage[SW_NST][KT + 2][64];
  double ybuf[2 * (KT + 1)][8];      // ring of the last KT+1 solved tile-row blocks, stored twice (no wrap-around arithmetic)
  double farpart[2][3][8];
  unsigned long long full[SW_NST];
};

// One directional sweep over tile rows.  DIR=+1: forward with Lb and the explicit D^-1 (rows r0..r1-1
// ascending), DIR=-1: backward with Ub, unit block diagonal (rows r1-1..r0 descending).
// Only blocks solved earlier in the same sweep contribute (the sweep range is the solve window).
// vin[row] is the right-hand side (streamed through the
// bulk-copy ring together with the factor tiles), sink(I, g, value) consumes the result.
// far warp FW (0..2) owns stage tiles FW, FW+3, ... of the far set: everything about them is a compile-time
// constant, the loop is straight-line code (4-5 x {tile LDS, y LDS, 2 FMA}) and two shuffles.
template <int KT, int DIR, int FW>
__device__ __forceinline__ double far_partial(const double* stage_row, const double* ybase, int lane, int tq) {
  double acc = 0.0, acc1 = 0.0;   // two accumulation chains
#pragma unroll
  for (int t = FW; t < KT - 1; t += 3) {
    // forward: stage tile tt (d = tt-KT) multiplies the block solved KT-tt iterations before the target row;
    // backward: stage tile tt (d = tt+1) multiplies the block solved tt+1 iterations before it
    const int tt = DIR > 0 ? t : t + 1;
    const int dist = DIR > 0 ? KT - tt : tt + 1;
    const double2 tv = *reinterpret_cast<const double2*>(stage_row + tt * 64 + 2 * lane);
    const double2 yp = *reinterpret_cast<const double2*>(ybase - dist * 8 + 2 * tq);
    acc = fma(tv.x, yp.x, acc);
    acc1 = fma(tv.y, yp.y, acc1);
  }
  acc += acc1;
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  return acc;
}

template <int KT, int SW_NST, int DIR, class Sink>
__device__ __forceinline__ void sweep_dir(SweepSmem<KT, SW_NST>& S, const SweepArgs& a, unsigned& itbase, int64_t r0, int64_t r1,
                                          const double* vin, int64_t nvalid, Sink sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int nrows = (int)(r1 - r0);
  if (nrows <= 0) return;
  constexpr int RING = KT + 1;
  // the ring's mbarriers keep counting across sweeps: global iteration index = itbase + it (32-bit, wraps)
  const unsigned ib = itbase;
  itbase += (unsigned)nrows;
  const bool bulk_rhs = ((reinterpret_cast<uintptr_t>(vin) & 15) == 0);
  // earlier generic-proxy writes of this CTA (previous sweep's results) must be visible to the async proxy
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  // y ring, written twice (slot and slot+RING) so that "the block solved d iterations ago" is a plain
  // subtraction; zero-filled: blocks outside the sweep window contribute nothing
  for (int e = threadIdx.x; e < 2 * RING * 8; e += blockDim.x) (&S.ybuf[0][0])[e] = 0.0;
  for (int e = threadIdx.x; e < 2 * 3 * 8; e += blockDim.x) (&S.farpart[0][0][0])[e] = 0.0;
  __syncthreads();
  constexpr int NTILE = DIR > 0 ? KT + 1 : KT;
  const int64_t rstart = DIR > 0 ? r0 : r1 - 1;                 // tile row of iteration 0
  const double* src0 = a.band + (rstart * a.tpr + (DIR > 0 ? 0 : KT + 1)) * SPK_TILE_ELEMS;  // d=-KT..0  or  d=1..KT
  const int64_t src_step = (int64_t)DIR * a.tpr * SPK_TILE_ELEMS;
  const double* rhs0 = vin + rstart * 8;
  const int64_t rows_bulk = (nvalid / 8);                       // tile rows whose 8 entries all exist in vin
  auto rhs_bulk_ok = [&](int it) -> bool { return bulk_rhs && (rstart + (int64_t)DIR * it) < rows_bulk; };
  auto issue = [&](int it) {  // executed by one thread
    const unsigned gi = ib + (unsigned)it;
    const int st = (int)(gi % SW_NST);
    uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[st]);
    const bool rb = rhs_bulk_ok(it);
    mbar_expect_tx(bar, (uint32_t)(NTILE * 512 + (rb ? 64 : 0)));
    bulk_g2s(&S.stage[st][0][0], src0 + (int64_t)it * src_step, NTILE * 512, bar);
    if (rb) bulk_g2s(&S.stage[st][KT + 1][0], rhs0 + (int64_t)DIR * it * 8, 64, bar);
  };
  if (threadIdx.x == SW_COPY_THREAD) {
    for (int it = 0; it < SW_NST && it < nrows; ++it) issue(it);
  }
  // iterations for which the right-hand side block came with the bulk copy: [bulk_lo, bulk_hi)
  // Every role runs its own loop (one CTA barrier per iteration closes all of them).
  if (warp == 0) {
    // -------- near warp: finishes the tile row of every iteration; the block solved one iteration ago never
    // leaves its registers (two shuffles issued before the barrier hand every lane its two entries)
    mbar_wait(reinterpret_cast<uint64_t*>(&S.full[ib % SW_NST]), (ib / SW_NST) & 1u);   // (the far warps wait for all later stages)
    int slot = 0;
    unsigned st = ib % SW_NST;
    double2 yp = make_double2(0.0, 0.0);
    for (int it = 0; it < nrows; ++it) {
#ifdef SW_TRACE
      const bool tr_on = blockIdx.x == SW_TRACE_CTA && DIR > 0 && it >= 32 && it < 64 && ib == 0;
#endif
      SW_TR(0);
      const int par = it & 1;
      const int64_t I = rstart + (int64_t)DIR * it;
      double rhs;
      if (rhs_bulk_ok(it)) rhs = S.stage[st][KT + 1][g];
      else rhs = (I * 8 + g < nvalid) ? vin[I * 8 + g] : 0.0;
      const double cg = rhs - (S.farpart[par][0][g] + S.farpart[par][1][g] + S.farpart[par][2][g]);
#ifdef SW_TRACE
      if (tr_on && cg == 1.2345e300) g_sw_trace[0][15] = 1;   // (dependence: the stamp below waits for cg)
#endif
      SW_TR(1);
      // adjacent tile: forward Lb(I,I-1) is stage tile KT-1 (d=-1); backward Ub(I,I+1) is stage tile 0 (d=+1)
#ifdef SW_TRACE_FINE   // stamps between the single operations of the chain (each pinned by an empty asm on its result)
      double2 t = *reinterpret_cast<const double2*>(&S.stage[st][DIR > 0 ? KT - 1 : 0][2 * lane]);
      asm volatile("" : "+d"(t.x), "+d"(t.y)); SW_TR(7);
      double part = fma(t.x, yp.x, t.y * yp.y);
      asm volatile("" : "+d"(part)); SW_TR(8);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      asm volatile("" : "+d"(part)); SW_TR(9);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      asm volatile("" : "+d"(part)); SW_TR(10);
      double yv = cg - part;
#else
      const double2 t = *reinterpret_cast<const double2*>(&S.stage[st][DIR > 0 ? KT - 1 : 0][2 * lane]);
      double part = fma(t.x, yp.x, t.y * yp.y);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      double yv = cg - part;  // replicated in the 4 lanes of row g
#endif
#ifdef SW_TRACE
      if (tr_on && yv == 1.2345e300) g_sw_trace[0][15] = 1;
#endif
      SW_TR(2);
      if (DIR > 0) {
        // y_g = sum_c Dinv[g][c] t_c
        const double2 dv = *reinterpret_cast<const double2*>(&S.stage[st][KT][2 * lane]);
        const double t0 = __shfl_sync(0xffffffffu, yv, 8 * tq);
        const double t1 = __shfl_sync(0xffffffffu, yv, 8 * tq + 4);
        yv = fma(dv.x, t0, dv.y * t1);
        yv += __shfl_xor_sync(0xffffffffu, yv, 1);
        yv += __shfl_xor_sync(0xffffffffu, yv, 2);
      }
#ifdef SW_TRACE
      if (tr_on && yv == 1.2345e300) g_sw_trace[0][15] = 1;
#endif
      SW_TR(3);
      if (tq == 0) { S.ybuf[slot][g] = yv; S.ybuf[slot + RING][g] = yv; }
      yp.x = __shfl_sync(0xffffffffu, yv, 8 * tq);
      yp.y = __shfl_sync(0xffffffffu, yv, 8 * tq + 4);
      SW_TR(4);
      __syncthreads();
      SW_TR(5);
      if (tq == 0) sink(I, g, yv);
      SW_TR(6);
      slot = slot + 1 == RING ? 0 : slot + 1;
      st = (st + 1) % SW_NST;
    }
  } else if (warp <= 3) {
    // -------- far warps: partial sums for the NEXT iteration from blocks solved >= 2 iterations before it
    int sn = 1 == RING ? 0 : 1;            // ring slot of iteration it+1
    unsigned gn = ib + 1u;
    for (int it = 0; it < nrows; ++it, ++gn) {
#ifdef SW_TRACE
#ifdef SW_TRACE_FINE
      const bool tr_on = false;
#else
      const bool tr_on = blockIdx.x == SW_TRACE_CTA && DIR > 0 && it >= 32 && it < 64 && ib == 0 && warp == 1;
#endif
#endif
      SW_TR(8);
      if (it + 1 < nrows) {
        const unsigned stn = gn % SW_NST;
        mbar_wait(reinterpret_cast<uint64_t*>(&S.full[stn]), (gn / SW_NST) & 1u);
        SW_TR(9);
        const double* srow = &S.stage[stn][0][0];
        const double* ybase = &S.ybuf[sn + RING][0];
        double acc;
        if (warp == 1) acc = far_partial<KT, DIR, 0>(srow, ybase, lane, tq);
        else if (warp == 2) acc = far_partial<KT, DIR, 1>(srow, ybase, lane, tq);
        else acc = far_partial<KT, DIR, 2>(srow, ybase, lane, tq);
        if (tq == 0) S.farpart[(it + 1) & 1][warp - 1][g] = acc;
#ifdef SW_TRACE
        if (tr_on && acc == 1.2345e300) g_sw_trace[0][15] = 1;
#endif
        SW_TR(10);
      }
      __syncthreads();
      SW_TR(11);
      sn = sn + 1 == RING ? 0 : sn + 1;
    }
  } else {
    // -------- copy warp: refill the stage the near warp finished one iteration ago
    for (int it = 0; it < nrows; ++it) {
#ifdef SW_TRACE
      const bool tr_on = blockIdx.x == SW_TRACE_CTA && DIR > 0 && it >= 32 && it < 64 && ib == 0;
#endif
      SW_TR(12);
      if (threadIdx.x == SW_COPY_THREAD && it >= 1 && it - 1 + SW_NST < nrows) issue(it - 1 + SW_NST);
      SW_TR(13);
      __syncthreads();
      SW_TR(14);
    }
  }
}

template <int KT, int SW_NST>
__global__ void __launch_bounds__(SW_THREADS, SW_NST > 6 ? 3 : 4) k_sweep(const SweepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SweepSmem<KT, SW_NST>& S = *reinterpret_cast<SweepSmem<KT, SW_NST>*>(smem_raw);
  const int kp = KT * 8;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SW_NST; ++i) mbar_init(reinterpret_cast<uint64_t*>(&S.full[i]), 1);
    fence_mbar_init();
  }
  pdl_trigger();
  __syncthreads();
  pdl_wait();   // everything above touched shared memory only
  unsigned itbase = 0;
  const int64_t n = a.n;
  if (a.mode == SWEEP_MAIN) {
    const int p = blockIdx.x;
    const int64_t t0 = a.pstart[p], t1 = a.pstart[p + 1];
    double* x = a.x;
    auto store_x = [&](int64_t I, int g, double v) { if (I * 8 + g < n) x[I * 8 + g] = v; };
    sweep_dir<KT, SW_NST, +1>(S, a, itbase, t0, t1, a.in, n, store_x);
    sweep_dir<KT, SW_NST, -1>(S, a, itbase, t0, t1, x, n, store_x);
    return;
  }
  // ---- corrections: blockIdx = 2*p + side (0 top, 1 bottom); when the window covers more than half
  //      of the partition both right-hand sides are folded into one full-partition job (side 0).
  //      w = A_i^{-1} r is built in the scratch vector, then x -= w over the window.
  const int p = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int64_t t0 = a.pstart[p], t1 = a.pstart[p + 1];
  const int64_t plen = t1 - t0;
  const bool top_on = (p > 0) || a.has_left;
  const bool bot_on = (p < a.P - 1) || a.has_right;
  const bool full = 2 * (int64_t)a.tipT > plen;
  const int64_t W = full ? plen : a.tipT;
  double* x = a.x;
  double* w = a.work;
  const double* rt = a.rtop + (size_t)p * kp;
  const double* rb = a.rbot + (size_t)p * kp;
  const int64_t npad = a.pstart[a.P] * 8;
  auto store_w = [&](int64_t I, int g, double v) { w[I * 8 + g] = v; };
  int64_t lo, hi, flo;   // window [lo,hi) ; forward sweep starts at flo (zero right-hand side above it)
  bool use_top, use_bot;
  if (full) {
    if (side == 1 || (!top_on && !bot_on)) return;
    lo = t0; hi = t1; use_top = top_on; use_bot = bot_on; flo = top_on ? t0 : t1 - KT;
  } else if (side == 0) {
    if (!top_on) return;
    lo = t0; hi = t0 + W; use_top = true; use_bot = false; flo = lo;
  } else {
    if (!bot_on) return;
    lo = t1 - W; hi = t1; use_top = false; use_bot = true; flo = t1 - KT;
  }
  for (int64_t e = lo * 8 + threadIdx.x; e < hi * 8; e += blockDim.x) {
    const int64_t I = e >> 3;
    double v = 0.0;
    if (use_top && I < t0 + KT) v += rt[e - t0 * 8];
    if (use_bot && I >= t1 - KT) v += rb[e - (t1 - KT) * 8];
    w[e] = v;
  }
  sweep_dir<KT, SW_NST, +1>(S, a, itbase, flo, hi, w, npad, store_w);
  sweep_dir<KT, SW_NST, -1>(S, a, itbase, lo, hi, w, npad, store_w);
  __syncthreads();
  for (int64_t e = lo * 8 + threadIdx.x; e < hi * 8; e += blockDim.x)
    if (e < n) x[e] -= w[e];
}

template <int KT, int NST>
static int launch_sweep_kt(spk_ctx* c, const SweepArgs& a, int grid) {
  const size_t smem = sizeof(SweepSmem<KT, NST>);
  SPK_CUDA(c, cudaFuncSetAttribute(k_sweep<KT, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // largest shared-memory carve-out: the CTAs of the spike-tip kernels (side stream, capi.cu) fit next to two resident sweep CTAs
  SPK_CUDA(c, cudaFuncSetAttribute(k_sweep<KT, NST>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  SPK_CUDA(c, spk_launch_pdl(k_sweep<KT, NST>, dim3(grid), dim3(SW_THREADS), smem, c->stream, a));
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}
static int launch_sweep_any(spk_ctx* c, const SweepArgs& a, int grid) {
  switch (c->L.kt) {
#define CASE(K_) case K_: return a.mode == SWEEP_MAIN ? launch_sweep_kt<K_, SW_NST_MAIN>(c, a, grid) : launch_sweep_kt<K_, SW_NST_CORR>(c, a, grid);
    CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    default: SPK_SET_ERR(c, "unsupported kt=%d", c->L.kt); return SPK_ERR_UNSUPPORTED;
  }
}

int spk_launch_sweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld) {
  if (c->wide) return spk_wide_main_sweep(c, b, x, nrhs, ld);
  for (int r = 0; r < nrhs; ++r) {
    SweepArgs a{};
    a.band = c->band; a.tpr = c->L.tpr; a.pstart = c->d_pstart; a.P = c->P;
    a.mode = SWEEP_MAIN; a.in = b + (size_t)r * ld; a.x = x + (size_t)r * ld; a.tipT = c->tipT; a.n = c->L.n;
    const int rc = launch_sweep_any(c, a, c->P);
    if (rc) return rc;
  }
  return SPK_OK;
}

int spk_launch_corrections(spk_ctx* c, double* x, int nrhs, int64_t ld) {
  if (c->wide) {
    for (int r = 0; r < nrhs; ++r) {
      const double* rt = c->gtip + (size_t)r * 2 * c->P * c->kp;
      const int rc = spk_wide_corrections(c, x + (size_t)r * ld, 1, ld, rt, rt + (size_t)c->P * c->kp, 0, c->work, c->L.nt * 8);
      if (rc) return rc;
    }
    return SPK_OK;
  }
  for (int r = 0; r < nrhs; ++r) {
    SweepArgs a{};
    a.band = c->band; a.tpr = c->L.tpr; a.pstart = c->d_pstart; a.P = c->P;
    a.mode = SWEEP_CORR; a.work = c->work; a.x = x + (size_t)r * ld; a.tipT = c->tipT; a.n = c->L.n;
    a.rtop = c->gtip + (size_t)r * 2 * c->P * c->kp;
    a.rbot = a.rtop + (size_t)c->P * c->kp;
    a.has_left = (c->opts.rank > 0); a.has_right = (c->opts.rank + 1 < c->opts.nranks);
    const int rc = launch_sweep_any(c, a, 2 * c->P);
    if (rc) return rc;
  }
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// reduced system solve + coupling right-hand sides, one CTA (32 warps: the loads of a whole block mat-vec in flight at once) per interface
// --------------------------------------------------------------------------------------------
struct RedSolveArgs {
  const double* band; BandLayout L; const int64_t* pstart;
  const double* Vb; const double* Wt; const double* Rinv;
  const double* x;          // g (after the main sweep)
  double* rtop; double* rbot;  // P * kp each
  int first_iface;
  int boundary_iface;          // interface index served with remote data (-1: none)
  const double* remoteWt; const double* remoteGtop; double* xbBoundary;
  int64_t n;
  int64_t x_stride; size_t tip_stride;   // several right-hand sides: blockIdx.y selects x + y*x_stride, rtop/rbot + y*tip_stride
  size_t bnd_stride;                     // ... and remoteGtop / xbBoundary + y*bnd_stride
};
// out[r] = base[r] - sum_c M[r*kp+c] v[c] (or just the product when base == nullptr).  8 warps; each warp
// takes 4 rows per pass and issues all their loads before the shuffle reductions (memory-level
// parallelism: the matrices stream from L2/HBM once per solve).
__device__ __forceinline__ void block_matvec(const double* __restrict__ M, const double* v, const double* base, double* out,
                                             int kp, bool negate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r0 = warp * 4; r0 < kp; r0 += nw * 4) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    double mv[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int q = 0; q < 4; ++q) { const int c = lane + 32 * q, r = r0 + m; mv[m][q] = (r < kp && c < kp) ? M[(size_t)r * kp + c] : 0.0; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      const double vc = (c < kp) ? v[c] : 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) s[m] = fma(mv[m][q], vc, s[m]);
    }
    for (int c = lane + 128; c < kp; c += 32) {   // kp > 128 never happens (kt <= 16), kept for safety
#pragma unroll
      for (int m = 0; m < 4; ++m) if (r0 + m < kp) s[m] = fma(M[(size_t)(r0 + m) * kp + c], v[c], s[m]);
    }
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int m = 0; m < 4; ++m) s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
    }
    if (lane < 4 && r0 + lane < kp) {
      const double sv = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      out[r0 + lane] = negate ? (base ? base[r0 + lane] : 0.0) - sv : sv;
    }
  }
}
__global__ void __launch_bounds__(1024) k_reduced_solve(const RedSolveArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int kp = a.L.kc * 8, KT = a.L.kc;
  double* gb = sm; double* gt = gb + kp; double* tv = gt + kp; double* xt = tv + kp; double* xb = xt + kp;
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x + a.first_iface;
  const bool bnd = (i == a.boundary_iface);
  const int64_t tb = a.pstart[i + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* xg = a.x + (size_t)blockIdx.y * a.x_stride;
  double* const rtop_o = a.rtop + (size_t)blockIdx.y * a.tip_stride;
  double* const rbot_o = a.rbot + (size_t)blockIdx.y * a.tip_stride;
  for (int e = threadIdx.x; e < kp; e += blockDim.x) {
    gb[e] = xg[(tb - KT) * 8 + e];
    gt[e] = bnd ? a.remoteGtop[(size_t)blockIdx.y * a.bnd_stride + e] : ((tb * 8 + e < a.n) ? xg[tb * 8 + e] : 0.0);
  }
  __syncthreads();
  const double* W = bnd ? a.remoteWt : a.Wt + (size_t)(i + 1) * kp * kp;
  const double* V = a.Vb + (size_t)i * kp * kp;
  const double* R = a.Rinv + (size_t)i * kp * kp;
  block_matvec(W, gb, gt, tv, kp, true);      // t  = g_t - W g_b
  __syncthreads();
  block_matvec(R, tv, nullptr, xt, kp, false); // x_t = R t
  __syncthreads();
  block_matvec(V, xt, gb, xb, kp, true);      // x_b = g_b - V x_t
  __syncthreads();
  // r_top of partition i+1: C_{i+1} x_b ;  r_bot of partition i: B_i x_t
  for (int r = warp; r < kp; r += (int)(blockDim.x >> 5)) {
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane; c < kp; c += 32) {
      if (!bnd && (c >> 3) >= (r >> 3)) s1 = fma(a.band[a.L.elem_off(tb * 8 + r, (tb - KT) * 8 + c)], xb[c], s1);
      if ((c >> 3) <= (r >> 3)) s2 = fma(a.band[a.L.elem_off((tb - KT) * 8 + r, tb * 8 + c)], xt[c], s2);
    }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) {
      if (!bnd) rtop_o[(size_t)(i + 1) * kp + r] = s1; else a.xbBoundary[(size_t)blockIdx.y * a.bnd_stride + r] = xb[r];
      rbot_o[(size_t)i * kp + r] = s2;
    }
  }
}

// r_top of partition 0 from the left neighbour's x_b:  r = C_0 x_b,  C_0(r,c) = A(r, c - kp); one CTA per column
__global__ void __launch_bounds__(1024) k_rtop_left(const double* __restrict__ band, BandLayout L, const double* __restrict__ xb_all, double* rtop_all,
                                                    size_t tip_stride) {
  const int kp = L.kc * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* xb = xb_all + (size_t)blockIdx.x * kp;
  double* rtop = rtop_all + (size_t)blockIdx.x * tip_stride;
  for (int r = warp; r < kp; r += (int)(blockDim.x >> 5)) {
    double s1 = 0.0;
    for (int c = lane; c < kp; c += 32)
      if ((c >> 3) >= (r >> 3)) s1 = fma(band[L.elem_off(r, (int64_t)c - kp)], xb[c], s1);
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if (lane == 0) rtop[r] = s1;
  }
}
int spk_launch_rtop_left(spk_ctx* c, double* rtop, size_t tip_stride, int nrhs) {
  k_rtop_left<<<nrhs, 1024, 0, c->stream>>>(c->band, c->L, c->remoteXbot, rtop, tip_stride);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

int spk_launch_reduced_solve(spk_ctx* c, double* x, int nrhs, int64_t ld, int iface_lo, int iface_hi) {
  if (c->wide) return spk_wide_reduced_solve(c, x, nrhs, ld, c->gtip, c->gtip + (size_t)c->P * c->kp, 2 * (size_t)c->P * c->kp);
  const bool has_right = c->opts.rank + 1 < c->opts.nranks;
  if (has_right && iface_hi == c->P - 1) iface_hi = c->P;   // include the boundary interface
  const int n = iface_hi - iface_lo;
  if (n <= 0) return SPK_OK;
  for (int r = 0; r < nrhs; ++r) {
    RedSolveArgs a;
    a.band = c->band; a.L = c->L; a.pstart = c->d_pstart; a.Vb = c->Vb; a.Wt = c->Wt; a.Rinv = c->Red;
    a.x = x + (size_t)r * ld;
    a.rtop = c->gtip + (size_t)r * 2 * c->P * c->kp;
    a.rbot = a.rtop + (size_t)c->P * c->kp;
    a.first_iface = iface_lo;
    a.boundary_iface = has_right ? c->P - 1 : -1;
    a.remoteWt = c->remoteWt; a.remoteGtop = c->remoteGtop; a.xbBoundary = c->xbBoundary; a.n = c->L.n;
    a.x_stride = 0; a.tip_stride = 0; a.bnd_stride = 0;
    // as many warps per CTA as keep every interface resident at once (2048 threads per SM): 32 warps = one pass per
    // block mat-vec when there are at most two interfaces per SM
    const int per_sm = (n + c->sm_count - 1) / c->sm_count;
    const int threads = std::max(256, std::min(1024, (2048 / std::max(per_sm, 1)) / 32 * 32));
    SPK_CUDA(c, spk_launch_pdl(k_reduced_solve, dim3(n), dim3(threads), sizeof(double) * 5 * c->kp, c->stream, a));
    SPK_KERNEL_CHECK(c);
  }
  return SPK_OK;
}

// All right-hand sides in one launch (grid.y = column): tips[r] = rtop | rbot of column r, 2*P*kp doubles each.
// On a sharded context the boundary interface (remote W^(t), g^(t) in; x^(b) out, kp doubles per column) rides along.
int spk_launch_reduced_solve_multi(spk_ctx* c, double* x, int nrhs, int64_t ld, double* tips) {
  if (c->wide) return spk_wide_reduced_solve(c, x, nrhs, ld, tips, tips + (size_t)c->P * c->kp, 2 * (size_t)c->P * c->kp);
  const bool has_right = c->opts.rank + 1 < c->opts.nranks;
  const int n = c->P - 1 + (has_right ? 1 : 0);
  if (n <= 0) return SPK_OK;
  RedSolveArgs a;
  a.band = c->band; a.L = c->L; a.pstart = c->d_pstart; a.Vb = c->Vb; a.Wt = c->Wt; a.Rinv = c->Red;
  a.x = x; a.rtop = tips; a.rbot = tips + (size_t)c->P * c->kp;
  a.first_iface = 0; a.boundary_iface = has_right ? c->P - 1 : -1;
  a.remoteWt = c->remoteWt; a.remoteGtop = c->remoteGtop; a.xbBoundary = c->xbBoundary; a.n = c->L.n;
  a.x_stride = ld; a.tip_stride = 2 * (size_t)c->P * c->kp; a.bnd_stride = (size_t)c->kp;
  for (int r0 = 0; r0 < nrhs; r0 += 65535) {
    const int nr = std::min(nrhs - r0, 65535);
    const int per_sm = (int)(((int64_t)n * nr + c->sm_count - 1) / c->sm_count);
    const int threads = std::max(256, std::min(1024, (2048 / std::max(per_sm, 1)) / 32 * 32));
    RedSolveArgs b = a;
    b.x = x + (size_t)r0 * ld; b.rtop = a.rtop + (size_t)r0 * a.tip_stride; b.rbot = a.rbot + (size_t)r0 * a.tip_stride;
    b.remoteGtop = a.remoteGtop + (size_t)r0 * a.bnd_stride; b.xbBoundary = a.xbBoundary + (size_t)r0 * a.bnd_stride;
    SPK_CUDA(c, spk_launch_pdl(k_reduced_solve, dim3(n, nr), dim3(threads), sizeof(double) * 5 * c->kp, c->stream, b));
    SPK_KERNEL_CHECK(c);
  }
  return SPK_OK;
}
