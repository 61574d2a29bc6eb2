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

struct WideSweepArgs { const WideSweepJob* jobs; int tpr, kts, KB; long long* trace; };
#define WS_ADD(slot) do { if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { const long long n_ = clock64(); tr[slot] += n_ - tmark; tmark = n_; } } while (0)

#define WS_THREADS 256
#define WS_WARPS 8

// one TMA-engine request pulls a whole 4 KB run (a tile row of a super-block) into L2
__device__ __forceinline__ void prefetch_l2_4k(const double* p) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], 4096;" ::"l"(p) : "memory"); }
__device__ __forceinline__ double2 ldnc_v2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// NSLOT = KB (window in super-blocks); warp r of the CTA owns tile row r of every super-block of the window: NSLOT
// accumulator tiles per right-hand-side tile column, slot of super-block J = J % NSLOT.  255 registers per thread
// (8 warps): the factor-tile ring (PFT tiles = 4*PFT registers) never spills -- a spilled ring slot would turn its
// load into a synchronous wait for HBM.
template <int NSLOT, int NCT, int PFT>
__global__ void __launch_bounds__(WS_THREADS, 1) k_wide_sweep(const WideSweepArgs a) {
  __shared__ __align__(16) double Cbuf[8][NCT][64];
  __shared__ __align__(16) double Ybuf[2][8][NCT][64];
  const WideSweepJob job = a.jobs[blockIdx.x];
  const int col0 = blockIdx.y * 8 * NCT;
  if (col0 >= job.ncols) return;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int tpr = a.tpr, kts = a.kts;
  constexpr int NIT = 8 * NSLOT;                       // factor tiles per warp and step
  constexpr int GS = (NSLOT % 4 == 0) ? 4 : 2;         // slots interleaved in the update loop (independent DMMA chains)
  static_assert(NIT % PFT == 0 && NSLOT % GS == 0, "ring depth must divide the tiles per step");
  const long long lo = job.sb_lo, hi = job.sb_hi;
  auto tile = [&](long long I, long long J) -> const double* { return job.band + (I * tpr + (J - I + kts)) * SPK_TILE_ELEMS; };
  // the super-block held in slot q when the window starts at super-block wb
  auto sb_of = [&](long long wb, int q) -> long long {
    long long m = wb % NSLOT;
    if (m < 0) m += NSLOT;
    long long d = q - m;
    if (d < 0) d += NSLOT;
    return wb + d;
  };
  // tile n of the update loop -> (slot q, k): groups of GS slots, k-major inside a group
  auto slot_of = [](int n) -> int { return (n / (8 * GS)) * GS + (n % GS); };
  auto k_of = [](int n) -> int { return (n % (8 * GS)) / GS; };
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
  double2 ring[PFT];
  long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = a.trace ? clock64() : 0;

  // =========================================== forward ===========================================
  {
    const long long I0 = job.sb_fwd;
    const bool ident = (job.in == nullptr);
    // window of step I0 before its update: super-block rows I0 .. I0+KB-1
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const long long sb = sb_of(I0, q);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[q][ct] = (sb < hi) ? rhs_pair(job.in, job.in_rs, job.in_cs, sb * 8 + warp, ct, ident) : make_double2(0.0, 0.0);
    }
    // factor tile n of step I: Lb(8*sb + warp, 8I + k), sb = the super-block in slot q of the window I+1 .. I+KB
    auto ftile = [&](long long I, int n) -> const double* {
      const long long sb = sb_of(I + 1, slot_of(n));
      return (I < hi && sb < hi) ? tile(sb * 8 + warp, I * 8 + k_of(n)) + 2 * lane : nullptr;
    };
#pragma unroll
    for (int n = 0; n < PFT; ++n) { const double* p = ftile(I0, n); ring[n] = p ? ldnc_v2(p) : make_double2(0.0, 0.0); }
    for (long long I = I0; I < hi; ++I) {
      const int par = (int)(I & 1);
      WS_ADD(3);
      // ---- phase 1: every warp hands its tile row of c_I over and takes its tile row of the entering row I+KB
      long long mq = I % NSLOT;
      if (mq < 0) mq += NSLOT;
      const int qs = (int)mq;
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        if (q == qs) {
#pragma unroll
          for (int ct = 0; ct < NCT; ++ct) {
            store_transposed(&Cbuf[warp][ct][0], acc[q][ct], g, tq);
            acc[q][ct] = (I + NSLOT < hi) ? rhs_pair(job.in, job.in_rs, job.in_cs, (I + NSLOT) * 8 + warp, ct, ident) : make_double2(0.0, 0.0);
          }
        }
      }
      // D_I^-1 row tile for phase 2 (requested before the barrier)
      double2 dv[8];
      {
        const double* dsrc = tile(I * 8 + warp, I * 8) + 2 * lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) dv[k] = ldnc_v2(dsrc + k * 64);
      }
      WS_ADD(0);
      __syncthreads();
      WS_ADD(1);
      // ---- phase 2: y_I = D_I^-1 c_I  (8 independent DMMA chains per column tile)
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
        store_transposed(&Ybuf[par][warp][ct][0], y, g, tq);
        out_pair(I * 8 + warp, ct, y);
      }
      WS_ADD(2);
      __syncthreads();
      WS_ADD(1);
      // ---- phase 3: c_J -= Lb(J,I) y_I for the window rows below
      // (the factor tiles and D^-1 of step I+2 are pulled into L2 now, one bulk prefetch per 4 KB run: the
      //  register ring and the D^-1 loads of the coming steps then pay L2 latency, not HBM latency)
      if (I + 2 < hi) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) {
          const long long sb = sb_of(I + 3, q);
          if (sb < hi && lane == q) prefetch_l2_4k(tile(sb * 8 + warp, (I + 2) * 8));
        }
        if (lane == NSLOT) prefetch_l2_4k(tile((I + 2) * 8 + warp, (I + 2) * 8));
      }
#pragma unroll
      for (int n0 = 0; n0 < NIT; n0 += GS) {
        double2 na[GS], yb[NCT];
#pragma unroll
        for (int qi = 0; qi < GS; ++qi) {
          const int n = n0 + qi;
          na[qi] = neg2(ring[n % PFT]);
          const int nn = n + PFT;   // refill the ring slot with the tile PFT positions ahead
          const double* p = (nn < NIT) ? ftile(I, nn) : ftile(I + 1, nn - NIT);
          ring[n % PFT] = p ? ldnc_v2(p) : make_double2(0.0, 0.0);
        }
        const int k = k_of(n0);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) yb[ct] = *reinterpret_cast<const double2*>(&Ybuf[par][k][ct][2 * lane]);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
          for (int qi = 0; qi < GS; ++qi) dmma884(acc[slot_of(n0 + qi)][ct].x, acc[slot_of(n0 + qi)][ct].y, na[qi].x, yb[ct].x);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
          for (int qi = 0; qi < GS; ++qi) dmma884(acc[slot_of(n0 + qi)][ct].x, acc[slot_of(n0 + qi)][ct].y, na[qi].y, yb[ct].y);
      }
    }
  }
  WS_ADD(3);
  __syncthreads();
  // =========================================== backward ==========================================
  {
    const long long I0 = hi - 1;
    // window of step I0 before its update: super-block rows I0-KB+1 .. I0, right-hand side = y (in `out`)
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const long long sb = sb_of(I0 + 1 - NSLOT, q);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[q][ct] = (sb >= lo) ? rhs_pair(job.out, job.out_rs, job.out_cs, sb * 8 + warp, ct, false) : make_double2(0.0, 0.0);
    }
    // factor tile n of step I: Ub(8*sb + warp, 8I + k), sb = the super-block in slot q of the window I-KB .. I-1
    auto ftile = [&](long long I, int n) -> const double* {
      const long long sb = sb_of(I - NSLOT, slot_of(n));
      return (I >= lo && sb >= lo) ? tile(sb * 8 + warp, I * 8 + k_of(n)) + 2 * lane : nullptr;
    };
#pragma unroll
    for (int n = 0; n < PFT; ++n) { const double* p = ftile(I0, n); ring[n] = p ? ldnc_v2(p) : make_double2(0.0, 0.0); }
    for (long long I = I0; I >= lo; --I) {
      const int par = (int)(I & 1);
      WS_ADD(6);
      long long mq = I % NSLOT;
      if (mq < 0) mq += NSLOT;
      const int qs = (int)mq;
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) {
        if (q == qs) {
#pragma unroll
          for (int ct = 0; ct < NCT; ++ct) {
            store_transposed(&Ybuf[par][warp][ct][0], acc[q][ct], g, tq);
            out_pair(I * 8 + warp, ct, acc[q][ct]);
            acc[q][ct] = (I - NSLOT >= lo) ? rhs_pair(job.out, job.out_rs, job.out_cs, (I - NSLOT) * 8 + warp, ct, false) : make_double2(0.0, 0.0);
          }
        }
      }
      WS_ADD(4);
      __syncthreads();
      WS_ADD(5);
      if (I - 2 >= lo) {
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) {
          const long long sb = sb_of(I - 2 - NSLOT, q);
          if (sb >= lo && lane == q) prefetch_l2_4k(tile(sb * 8 + warp, (I - 2) * 8));
        }
      }
#pragma unroll
      for (int n0 = 0; n0 < NIT; n0 += GS) {
        double2 na[GS], yb[NCT];
#pragma unroll
        for (int qi = 0; qi < GS; ++qi) {
          const int n = n0 + qi;
          na[qi] = neg2(ring[n % PFT]);
          const int nn = n + PFT;
          const double* p = (nn < NIT) ? ftile(I, nn) : ftile(I - 1, nn - NIT);
          ring[n % PFT] = p ? ldnc_v2(p) : make_double2(0.0, 0.0);
        }
        const int k = k_of(n0);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) yb[ct] = *reinterpret_cast<const double2*>(&Ybuf[par][k][ct][2 * lane]);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
          for (int qi = 0; qi < GS; ++qi) dmma884(acc[slot_of(n0 + qi)][ct].x, acc[slot_of(n0 + qi)][ct].y, na[qi].x, yb[ct].x);
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct)
#pragma unroll
          for (int qi = 0; qi < GS; ++qi) dmma884(acc[slot_of(n0 + qi)][ct].x, acc[slot_of(n0 + qi)][ct].y, na[qi].y, yb[ct].y);
      }
    }
  }
  WS_ADD(6);
  if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) for (int q = 0; q < 8; ++q) a.trace[q] = tr[q];
}

template <int NSLOT, int NCT, int PFT>
static int launch_ws(spk_ctx* c, const WideSweepArgs& a, int njobs, int groups) {
  k_wide_sweep<NSLOT, NCT, PFT><<<dim3(njobs, groups), WS_THREADS, 0, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// run the jobs (device array); max_cols = the largest ncols among them
int spk_wide_sweep(spk_ctx* c, const WideSweepJob* d_jobs, int njobs, int max_cols) {
  if (njobs <= 0 || max_cols <= 0) return SPK_OK;
  WideSweepArgs a; a.jobs = d_jobs; a.tpr = c->L.tpr; a.kts = c->L.kt; a.KB = c->kb;
  a.trace = (c->lu_trace && d_jobs == (const WideSweepJob*)c->d_wjobs + (size_t)3 * c->wjobs_cap) ? (long long*)c->lu_trace + 128 : nullptr;   // (debug: the partition sweeps)
  const int nslot = c->kb;
  // 16 columns per CTA (the band is streamed once per 16 columns) unless that leaves most of the GPU idle: few jobs
  // (partitions) -> 8 columns per CTA, twice the CTAs, the second reader of a factor tile hits L2
  const bool one = max_cols <= 8 || (int64_t)njobs * ((max_cols + 15) / 16) * 4 < (int64_t)c->sm_count * 3;
  const int groups = one ? (max_cols + 7) / 8 : (max_cols + 15) / 16;
  for (int j0 = 0; j0 < njobs; j0 += 65535) {
    WideSweepArgs b = a; b.jobs = d_jobs + j0;
    const int nj = std::min(njobs - j0, 65535);
    int rc;
    switch (nslot * 2 + (one ? 0 : 1)) {
      case 8: rc = launch_ws<4, 1, 16>(c, b, nj, groups); break;
      case 9: rc = launch_ws<4, 2, 16>(c, b, nj, groups); break;
      case 12: rc = launch_ws<6, 1, 16>(c, b, nj, groups); break;
      case 13: rc = launch_ws<6, 2, 16>(c, b, nj, groups); break;
      case 16: rc = launch_ws<8, 1, 16>(c, b, nj, groups); break;
      case 17: rc = launch_ws<8, 2, 16>(c, b, nj, groups); break;
      default: SPK_SET_ERR(c, "wide sweep: unsupported window of %d super-blocks", c->kb); return SPK_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
  }
  return SPK_OK;
}
