// wide_sweep.cu -- block triangular sweeps for the wide-band factor format (wide.cuh), any number of right-hand
// sides, plus the small data-movement kernels of the wide path (coupling blocks, reversed tip windows, reduced
// matrices).  Replaces PCApply(inner) = MatSolve_SeqAIJ (/root/reference/src/matbanded.c:190) for K = 129..512 and
// is also the engine behind the wide spike tips (V^(b), W^(t) are sweeps with K right-hand sides) and the
// inverse of the reduced blocks.
//
// Right-looking ("column oriented") sweeps in super-block steps:  forward  y_I = D_I^-1 c_I, then
// c_J -= Lb(J,I) y_I for the KB super-block rows below;  backward  x_I = c_I, then c_J -= Ub(J,I) x_I for the KB
// rows above.  One CTA per (job, group of 8*NCT right-hand sides), 8 warps:
//   * the pending right-hand-side window (KB super-blocks = 8*KB tile rows x 8*NCT columns) lives in REGISTERS as
//     DMMA accumulator tiles: warp r holds tile row r of every super-block of the window, slot = super-block % KB;
//   * the factor tiles are the LEFT operands: they go from HBM/L2 straight into registers (row-major tile =
//     fragment), through a software ring PFT tiles deep per warp (register resident: 128 KB in flight per SM), no shared
//     memory staging; every factor entry is read exactly once per column group;
//   * the freshly solved block (64 x 8*NCT) is the RIGHT operand: it is broadcast through shared memory as
//     transposed tiles; one CTA barrier per step (two in the forward sweep, where D_I^-1 is applied in between).
#include "wide.cuh"
#include <algorithm>

struct WideSweepArgs { const WideSweepJob* jobs; int tpr, kts, KB; long long* trace; const double* zero; };
#define WS_ADD(slot) do { if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { const long long n_ = clock64(); tr[slot] += n_ - tmark; tmark = n_; } } while (0)

#define WS_THREADS 256
#ifndef WS_PFA
#define WS_PFA 0   // L2 bulk prefetch of the sweeps' factor runs, in super-block steps ahead; 0 = none (measured: 2 costs 8 %, r02_summary)
#endif
#define WS_WARPS 8

// one TMA-engine request pulls a whole 4 KB run (a tile row of a super-block) into L2
__device__ __forceinline__ void prefetch_l2_4k(const double* p) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], 4096;" ::"l"(p) : "memory"); }
__device__ __forceinline__ double2 ldnc_v2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// NSLOT = KB (window in super-blocks); warp r of the CTA owns tile row r of every super-block of the window: NSLOT
// accumulator tiles per right-hand-side tile column, acc[d] = the rows of super-block I+d (forward) / I-d (backward)
// at step I; the window slides by register renaming / moves.
// Addressing is affine: with P(I) = the warp's tile row of the DIAGONAL super-block of step I and DS = 512 (tpr - 1),
//     tile k of  Lb(I+d, I) = P(I) + d DS + 64 k,     Ub(I-d, I) = P(I) - d DS + 64 k,     D_I^-1 = P(I) + 64 k,
// so a step keeps one base pointer per window row (a zero run replaces rows outside the job) and every factor tile is
// ONE load with an immediate offset.  (The first version indexed the window by super-block % NSLOT: ~25 integer
// instructions per tile -- 64-bit multiplies, slot arithmetic, null-pointer branches -- made the kernel issue bound at
// 3x its DMMA time; see profiles/r02_summary.md.)  The factor tiles of the coming PFT positions sit in a register ring
// (255 registers per thread, 8 warps; a spilled ring slot would turn its load into a synchronous wait for HBM).
template <int NSLOT, int NCT, int PFT, bool SR>
__global__ void __launch_bounds__(WS_THREADS, 1) k_wide_sweep(const WideSweepArgs a) {
  __shared__ __align__(16) double Cbuf[8][NCT][64];
  __shared__ __align__(16) double Ybuf[2][8][NCT][64];
  extern __shared__ __align__(128) double ws_ring[];   // [8 warps][PFT tiles][64]: the factor-tile ring, filled by cp.async
  const WideSweepJob job = a.jobs[blockIdx.x];
  const int col0 = blockIdx.y * 8 * NCT;
  if (col0 >= job.ncols) return;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int tpr = a.tpr, kts = a.kts;
  constexpr int NIT = 8 * NSLOT;                       // factor tiles per warp and step
  constexpr int GS = (NSLOT % 4 == 0) ? 4 : 2;         // window rows interleaved in the update loop (independent DMMA chains)
  static_assert(NIT % PFT == 0 && NSLOT % GS == 0 && PFT % GS == 0, "ring depth must divide the tiles per step");
  static_assert(PFT <= NIT, "a refill reaches at most into the next step");
  const long long lo = job.sb_lo, hi = job.sb_hi;
  // tile n of the update loop -> (window row d = 1..NSLOT, k): groups of GS rows, k-major inside a group
  auto d_of = [](int n) -> int { return (n / (8 * GS)) * GS + (n % GS) + 1; };
  auto k_of = [](int n) -> int { return (n % (8 * GS)) / GS; };
  const long long DS = 512ll * (tpr - 1);              // doubles between a window row's run and the next one's
  const long long STEP = 512ll * tpr;                  // doubles between P(I) and P(I+1)
  const double* const zrun = a.zero + 2 * lane;        // 4 KB of zeros: the run of a window row outside the job
  auto diag_run = [&](long long I) -> const double* { return job.band + ((I * 8 + warp) * (long long)tpr + (kts - warp)) * SPK_TILE_ELEMS + 2 * lane; };
  auto rhs_pair = [&](const double* src, long long rs, long long cs, long long t, int ct, bool identity) -> double2 {
    const long long r = t * 8 + g - job.row0;
    const int c = col0 + ct * 8 + 2 * tq;
    double2 v = make_double2(0.0, 0.0);
    if (r >= 0 && r < job.nrow_valid) {
      if (identity) { v.x = (r == c) ? 1.0 : 0.0; v.y = (r == c + 1) ? 1.0 : 0.0; }
      else {
        if (c < job.ncols) v.x = src[r * rs + c * cs];
        if (c + 1 < job.ncols) v.y = src[r * rs + (c + 1) * cs];
      }
    }
    return v;
  };
  auto out_pair = [&](long long t, int ct, const double2& v) {
    const long long r = t * 8 + g - job.row0;
    const int c = col0 + ct * 8 + 2 * tq;
    if (r >= 0 && r < job.nrow_valid) {
      if (c < job.ncols) job.out[r * job.out_rs + c * job.out_cs] = v.x;
      if (c + 1 < job.ncols) job.out[r * job.out_rs + (c + 1) * job.out_cs] = v.y;
    }
  };

  double2 acc[NSLOT][NCT];
  // The factor-tile ring of a warp, PFT tiles deep, in one of two places (SR):
  //  * SHARED memory, private to the warp: lane l copies bytes [32 l, 32 l + 16) of a tile with cp.async and reads the same
  //    bytes back, so a cp.async.wait_group is all the synchronisation there is; 32-48 tiles in flight per warp without
  //    holding a register.  For the long jobs (partition sweeps, corrections), where a 16-tile register ring left the kernel
  //    waiting on the long scoreboard: 3.84 -> 2.77 ms at C5.
  //  * REGISTERS (16 tiles): for the many short jobs of the spike tips and reduced inverses (8-36 steps, 32 column groups
  //    per job), where the deeper ring's fill and drain per sweep cost more than its depth buys (0.63 vs 0.86 ms).
  double2 ring[SR ? 1 : PFT];
  double* const rbase = ws_ring + (size_t)warp * PFT * 64 + 2 * lane;
  const uint32_t rbase_s = smem_u32(rbase);
  constexpr int NG = PFT / GS;   // groups in flight
  auto ring_fill = [&](int slot, const double* src) {
    if (SR) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rbase_s + (uint32_t)slot * 512u), "l"(src) : "memory");
    else ring[SR ? 0 : slot] = ldnc_v2(src);
  };
  auto ring_commit = [&]() { if (SR) asm volatile("cp.async.commit_group;" ::: "memory"); };
  const double* pd[NSLOT];   // run of window row d+1 at the current step
  const double* pn[NSLOT];   // ... at the next step
  long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = a.trace ? clock64() : 0;

  // the update loop of one step, shared by both directions: acc[d-1] <- acc[d] + (factor tile row d) * Ybuf (which holds
  // the NEGATED solved block), d = 1..NSLOT with acc[NSLOT] = the entering rows `ent`
  auto update = [&](const double2 (&ent)[NCT], int par) {
    double2 nxt[NSLOT][NCT];
#pragma unroll
    for (int d = 1; d <= NSLOT; ++d)
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) nxt[d - 1][ct] = (d < NSLOT) ? acc[d < NSLOT ? d : 0][ct] : ent[ct];
#pragma unroll
    for (int n0 = 0; n0 < NIT; n0 += GS) {
      double2 fa[GS], yb[NCT];
      if (SR) asm volatile("cp.async.wait_group %0;" ::"n"(NG - 1) : "memory");   // the group of this iteration's tiles has landed
#pragma unroll
      for (int qi = 0; qi < GS; ++qi) {
        const int n = n0 + qi, nn = n + PFT;
        if (SR) fa[qi] = *reinterpret_cast<const double2*>(rbase + (n % PFT) * 64);
        else {   // register ring: the slot is renamed into the operand and refilled at once
          fa[qi] = ring[SR ? 0 : n % PFT];
          ring_fill(n % PFT, (nn < NIT) ? pd[d_of(nn) - 1] + 64 * k_of(nn) : pn[d_of(nn - NIT) - 1] + 64 * k_of(nn - NIT));
        }
      }
      const int k = k_of(n0);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) yb[ct] = *reinterpret_cast<const double2*>(&Ybuf[par][k][ct][2 * lane]);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
        for (int qi = 0; qi < GS; ++qi) dmma884(nxt[d_of(n0 + qi) - 1][ct].x, nxt[d_of(n0 + qi) - 1][ct].y, fa[qi].x, yb[ct].x);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
        for (int qi = 0; qi < GS; ++qi) dmma884(nxt[d_of(n0 + qi) - 1][ct].x, nxt[d_of(n0 + qi) - 1][ct].y, fa[qi].y, yb[ct].y);
      // refill the slots just consumed (the DMMAs above have read them) with the tiles PFT positions ahead, possibly of
      // the next step; one group per iteration
      if (SR) {
#pragma unroll
        for (int qi = 0; qi < GS; ++qi) {
          const int n = n0 + qi, nn = n + PFT;
          ring_fill(n % PFT, (nn < NIT) ? pd[d_of(nn) - 1] + 64 * k_of(nn) : pn[d_of(nn - NIT) - 1] + 64 * k_of(nn - NIT));
        }
        ring_commit();
      }
    }
#pragma unroll
    for (int d = 0; d < NSLOT; ++d)
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[d][ct] = nxt[d][ct];
  };

  // =========================================== forward ===========================================
  {
    const long long I0 = job.sb_fwd;
    const bool ident = (job.in == nullptr);
    // window of step I0: super-block rows I0 .. I0+KB-1
#pragma unroll
    for (int d = 0; d < NSLOT; ++d)
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[d][ct] = (I0 + d < hi) ? rhs_pair(job.in, job.in_rs, job.in_cs, (I0 + d) * 8 + warp, ct, ident) : make_double2(0.0, 0.0);
    // runs of Lb(I+d, I), d = 1..NSLOT, for step I (zero run when the row or the step is outside the job)
    auto set_runs = [&](const double* (&pp)[NSLOT], long long I) {
      const double* P = diag_run(I);
#pragma unroll
      for (int d = 1; d <= NSLOT; ++d) pp[d - 1] = (I < hi && I + d < hi) ? P + d * DS : zrun;
    };
    set_runs(pd, I0);
    set_runs(pn, I0 + 1);
#pragma unroll
    for (int n = 0; n < PFT; ++n) {
      ring_fill(n, pd[d_of(n) - 1] + 64 * k_of(n));
      if (n % GS == GS - 1) ring_commit();
    }
    double2 dv[8];
    {
      const double* dsrc = (I0 < hi) ? diag_run(I0) : zrun;
#pragma unroll
      for (int k = 0; k < 8; ++k) dv[k] = ldnc_v2(dsrc + 64 * k);
    }
    for (long long I = I0; I < hi; ++I) {
      const int par = (int)(I & 1);
      WS_ADD(3);
      // ---- phase 1: every warp hands its tile row of c_I over
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) store_transposed(&Cbuf[warp][ct][0], acc[0][ct], g, tq);
      WS_ADD(0);
      __syncthreads();
      WS_ADD(1);
      // ---- phase 2: y_I = D_I^-1 c_I  (8 independent DMMA chains per column tile); the NEGATED block is what the
      //      updates multiply by
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) {
        double2 yk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) yk[k] = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const double2 cb = *reinterpret_cast<const double2*>(&Cbuf[k][ct][2 * lane]); dmma884(yk[k].x, yk[k].y, dv[k].x, cb.x); }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const double2 cb = *reinterpret_cast<const double2*>(&Cbuf[k][ct][2 * lane]); dmma884(yk[k].x, yk[k].y, dv[k].y, cb.y); }
        const double2 y = make_double2(((yk[0].x + yk[1].x) + (yk[2].x + yk[3].x)) + ((yk[4].x + yk[5].x) + (yk[6].x + yk[7].x)),
                                       ((yk[0].y + yk[1].y) + (yk[2].y + yk[3].y)) + ((yk[4].y + yk[5].y) + (yk[6].y + yk[7].y)));
        store_transposed(&Ybuf[par][warp][ct][0], neg2(y), g, tq);
        out_pair(I * 8 + warp, ct, y);
      }
      // D^-1 of the next step: requested a whole update phase ahead of its use
      {
        const double* dsrc = (I + 1 < hi) ? diag_run(I + 1) : zrun;
#pragma unroll
        for (int k = 0; k < 8; ++k) dv[k] = ldnc_v2(dsrc + 64 * k);
      }
      WS_ADD(2);
      __syncthreads();
      WS_ADD(1);
      // ---- phase 3: c_J -= Lb(J,I) y_I for the window rows below; row I+KB enters
#if WS_PFA > 0
      if (I + WS_PFA < hi) {   // factor tiles and D^-1 of a later step pulled into L2, one bulk prefetch per 4 KB run
        const double* P = diag_run(I + WS_PFA) - 2 * lane;
        if (lane <= NSLOT && I + WS_PFA + lane < hi) prefetch_l2_4k(P + lane * DS);
      }
#endif
      double2 ent[NCT];
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) ent[ct] = (I + NSLOT < hi) ? rhs_pair(job.in, job.in_rs, job.in_cs, (I + NSLOT) * 8 + warp, ct, ident) : make_double2(0.0, 0.0);
      update(ent, par);
#pragma unroll
      for (int d = 0; d < NSLOT; ++d) pd[d] = pn[d];
      set_runs(pn, I + 2);
    }
  }
  WS_ADD(3);
  if (SR) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // =========================================== backward ==========================================
  {
    const long long I0 = hi - 1;
    // window of step I0: super-block rows I0, I0-1, .. I0-KB+1, right-hand side = y (in `out`)
#pragma unroll
    for (int d = 0; d < NSLOT; ++d)
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[d][ct] = (I0 - d >= lo) ? rhs_pair(job.out, job.out_rs, job.out_cs, (I0 - d) * 8 + warp, ct, false) : make_double2(0.0, 0.0);
    // runs of Ub(I-d, I), d = 1..NSLOT: row I-d, i.e. the diagonal run of step I-d moved 8d tiles to the right
    auto set_runs = [&](const double* (&pp)[NSLOT], long long I) {
      const double* P = diag_run(I);
#pragma unroll
      for (int d = 1; d <= NSLOT; ++d) pp[d - 1] = (I >= lo && I - d >= lo) ? P - d * DS : zrun;
    };
    set_runs(pd, I0);
    set_runs(pn, I0 - 1);
#pragma unroll
    for (int n = 0; n < PFT; ++n) {
      ring_fill(n, pd[d_of(n) - 1] + 64 * k_of(n));
      if (n % GS == GS - 1) ring_commit();
    }
    for (long long I = I0; I >= lo; --I) {
      const int par = (int)(I & 1);
      WS_ADD(6);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) {
        store_transposed(&Ybuf[par][warp][ct][0], neg2(acc[0][ct]), g, tq);
        out_pair(I * 8 + warp, ct, acc[0][ct]);
      }
      WS_ADD(4);
      __syncthreads();
      WS_ADD(5);
#if WS_PFA > 0
      if (I - WS_PFA >= lo) {
        const double* P = diag_run(I - WS_PFA) - 2 * lane;
        if (lane >= 1 && lane <= NSLOT && I - WS_PFA - lane >= lo) prefetch_l2_4k(P - lane * DS);
      }
#endif
      double2 ent[NCT];
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) ent[ct] = (I - NSLOT >= lo) ? rhs_pair(job.out, job.out_rs, job.out_cs, (I - NSLOT) * 8 + warp, ct, false) : make_double2(0.0, 0.0);
      update(ent, par);
#pragma unroll
      for (int d = 0; d < NSLOT; ++d) pd[d] = pn[d];
      set_runs(pn, I - 2);
    }
  }
  WS_ADD(6);
  if (SR) asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) for (int q = 0; q < 8; ++q) a.trace[q] = tr[q];
}

template <int NSLOT, int NCT, int PFT, bool SR>
static int launch_ws2(spk_ctx* c, const WideSweepArgs& a, int njobs, int groups) {
  const size_t smem = SR ? sizeof(double) * (size_t)WS_WARPS * PFT * 64 : 0;
  if (SR) SPK_CUDA(c, cudaFuncSetAttribute(k_wide_sweep<NSLOT, NCT, PFT, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_wide_sweep<NSLOT, NCT, PFT, SR><<<dim3(njobs, groups), WS_THREADS, smem, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}
// long jobs: shared-memory ring, as deep as a step allows (a refill reaches at most into the next step); short jobs: 16 registers-tiles
template <int NSLOT, int NCT>
static int launch_ws(spk_ctx* c, const WideSweepArgs& a, int njobs, int groups, bool long_jobs) {
  constexpr int DEEP = (8 * NSLOT <= 48) ? 8 * NSLOT : 32;
  return long_jobs ? launch_ws2<NSLOT, NCT, DEEP, true>(c, a, njobs, groups) : launch_ws2<NSLOT, NCT, 16, false>(c, a, njobs, groups);
}

// run the jobs (device array); max_cols = the largest ncols among them
int spk_wide_sweep(spk_ctx* c, const WideSweepJob* d_jobs, int njobs, int max_cols, bool long_jobs) {
  if (njobs <= 0 || max_cols <= 0) return SPK_OK;
  WideSweepArgs a; a.jobs = d_jobs; a.tpr = c->L.tpr; a.kts = c->L.kt; a.KB = c->kb; a.zero = c->wide_zero;
  a.trace = (c->lu_trace && d_jobs == (const WideSweepJob*)c->d_wjobs + (size_t)3 * c->wjobs_cap) ? (long long*)c->lu_trace + 128 : nullptr;   // (debug: the partition sweeps)
  const int nslot = c->kb;
  // 16 columns per CTA (the band is streamed once per 16 columns) unless that leaves most of the GPU idle: few jobs
  // (partitions) -> 8 columns per CTA, twice the CTAs, the second reader of a factor tile hits L2
  bool one = max_cols <= 8 || (int64_t)njobs * ((max_cols + 15) / 16) * 4 < (int64_t)c->sm_count * 3;
  if (const char* e = getenv("SPIKE_WS_NCT")) one = max_cols <= 8 || atoi(e) == 1;   // tools: force 8 / 16 columns per CTA
  const int groups = one ? (max_cols + 7) / 8 : (max_cols + 15) / 16;
  for (int j0 = 0; j0 < njobs; j0 += 65535) {
    WideSweepArgs b = a; b.jobs = d_jobs + j0;
    const int nj = std::min(njobs - j0, 65535);
    int rc;
    switch (nslot * 2 + (one ? 0 : 1)) {
      case 8: rc = launch_ws<4, 1>(c, b, nj, groups, long_jobs); break;
      case 9: rc = launch_ws<4, 2>(c, b, nj, groups, long_jobs); break;
      case 12: rc = launch_ws<6, 1>(c, b, nj, groups, long_jobs); break;
      case 13: rc = launch_ws<6, 2>(c, b, nj, groups, long_jobs); break;
      case 16: rc = launch_ws<8, 1>(c, b, nj, groups, long_jobs); break;
      case 17: rc = launch_ws<8, 2>(c, b, nj, groups, long_jobs); break;
      default: SPK_SET_ERR(c, "wide sweep: unsupported window of %d super-blocks", c->kb); return SPK_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
  }
  return SPK_OK;
}
