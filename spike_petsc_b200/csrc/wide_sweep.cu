// wide_sweep.cu -- block triangular sweeps for the wide-band factor format (wide.cuh), any number of right-hand
// sides, plus the small data-movement kernels of the wide path (coupling blocks, reversed tip windows, reduced
// matrices).  Replaces PCApply(inner) = MatSolve_SeqAIJ (/root/reference/src/matbanded.c:190) for K = 129..512 and
// is also the engine behind the wide spike tips (V^(b), W^(t) are sweeps with K right-hand sides) and the
// inverse of the reduced blocks.
//
// Right-looking ("column oriented") sweeps in super-block steps:  forward  y_I = D_I^-1 c_I, then
// c_J -= Lb(J,I) y_I for the KB super-block rows below;  backward  x_I = c_I, then c_J -= Ub(J,I) x_I for the KB
// rows above.  One CTA per (job, group of 8*NCT right-hand sides), 16 warps:
//   * the pending right-hand-side window (KB super-blocks = 8*KB tile rows x 8*NCT columns) lives in REGISTERS as
//     DMMA accumulator tiles: tile row t belongs to warp t % 16, slot (t/16) % NSLOT (KB even, NSLOT = KB/2);
//   * the factor tiles are the LEFT operands: they go from HBM/L2 straight into registers (row-major tile =
//     fragment), through a software ring PFT tiles deep per warp -- up to 128 KB in flight per SM, no shared
//     memory staging; every factor entry is read exactly once per column group;
//   * the freshly solved block (64 x 8*NCT) is the RIGHT operand: it is broadcast through shared memory as
//     transposed tiles; one CTA barrier per step (two in the forward sweep, where D_I^-1 is applied in between).
#include "wide.cuh"
#include <algorithm>

struct WideSweepArgs { const WideSweepJob* jobs; int tpr, kts, KB; };

#define WS_THREADS 512
#define WS_WARPS 16

__device__ __forceinline__ double2 ldnc_v2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <int NSLOT, int NCT, int PFT>
__global__ void __launch_bounds__(WS_THREADS, 1) k_wide_sweep(const WideSweepArgs a) {
  __shared__ __align__(16) double Cbuf[8][NCT][64];
  __shared__ __align__(16) double Ybuf[2][8][NCT][64];
  const WideSweepJob job = a.jobs[blockIdx.x];
  const int col0 = blockIdx.y * 8 * NCT;
  if (col0 >= job.ncols) return;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int tpr = a.tpr, kts = a.kts, KB = a.KB;
  constexpr int WIN = 16 * NSLOT;   // tile rows in the window (= 8*KB)
  constexpr int NIT = 8 * NSLOT;    // factor tiles per warp and step
  static_assert(NIT % PFT == 0, "ring depth must divide the tiles per step");
  const long long lo8 = job.sb_lo * 8, hi8 = job.sb_hi * 8;
  auto tile = [&](long long I, long long J) -> const double* { return job.band + (I * tpr + (J - I + kts)) * SPK_TILE_ELEMS; };
  // the tile row held in (this warp, slot q) when the window starts at tile row wb (multiple of 8)
  auto tile_of = [&](long long wb, int q) -> long long {
    long long m = wb % WIN;
    if (m < 0) m += WIN;
    long long t = wb - m + q * 16 + warp;
    if (t < wb) t += WIN;
    return t;
  };
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

  // =========================================== forward ===========================================
  {
    const long long I0 = job.sb_fwd;
    const bool ident = (job.in == nullptr);
    // window of step I0 before its update: super-block rows I0 .. I0+KB-1
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const long long t = tile_of(I0 * 8, q);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[q][ct] = (t < hi8) ? rhs_pair(job.in, job.in_rs, job.in_cs, t, ct, ident) : make_double2(0.0, 0.0);
    }
    // factor tile n of step I: slot q = n/8, k = n%8 -> Lb(tile_of(8(I+1), q), 8I + k)
    auto ftile = [&](long long I, int n) -> const double* {
      const long long t = tile_of((I + 1) * 8, n >> 3);
      return (I < job.sb_hi && t < hi8) ? tile(t, I * 8 + (n & 7)) + 2 * lane : nullptr;
    };
#pragma unroll
    for (int n = 0; n < PFT; ++n) { const double* p = ftile(I0, n); ring[n] = p ? ldnc_v2(p) : make_double2(0.0, 0.0); }
    for (long long I = I0; I < job.sb_hi; ++I) {
      const int par = (int)(I & 1);
      // ---- phase 1: the owners of super-block row I hand c_I over and take the entering row I+KB
      const int rr = (int)((warp - (int)((I * 8) % 16) + 16) % 16);   // my tile row inside super-block I (if < 8)
      if (rr < 8) {
        const long long t = I * 8 + rr;
        const int qs = (int)((t / 16) % NSLOT);
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) {
          if (q == qs) {
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
              store_transposed(&Cbuf[rr][ct][0], acc[q][ct], g, tq);
              acc[q][ct] = (t + WIN < hi8) ? rhs_pair(job.in, job.in_rs, job.in_cs, t + WIN, ct, ident) : make_double2(0.0, 0.0);
            }
          }
        }
      }
      // D_I^-1 row tile for phase 2 (requested before the barrier)
      const int r2 = warp & 7;
      const bool p2 = (NCT == 2) || (warp < 8);
      const int ct2 = (NCT == 2) ? (warp >> 3) : 0;
      double2 dv[8];
      if (p2) {
        const double* dsrc = tile(I * 8 + r2, I * 8) + 2 * lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) dv[k] = ldnc_v2(dsrc + k * 64);
      }
      __syncthreads();
      // ---- phase 2: y_I = D_I^-1 c_I
      if (p2) {
        double2 y = make_double2(0.0, 0.0), y2 = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          dmma_cc(y, dv[k], *reinterpret_cast<const double2*>(&Cbuf[k][ct2][2 * lane]));
          dmma_cc(y2, dv[k + 1], *reinterpret_cast<const double2*>(&Cbuf[k + 1][ct2][2 * lane]));
        }
        y.x += y2.x; y.y += y2.y;
        store_transposed(&Ybuf[par][r2][ct2][0], y, g, tq);
        out_pair(I * 8 + r2, ct2, y);
      }
      __syncthreads();
      // ---- phase 3: c_J -= Lb(J,I) y_I for the window rows below
#pragma unroll
      for (int n = 0; n < NIT; ++n) {
        const int q = n >> 3, k = n & 7;
        const double2 na = neg2(ring[n % PFT]);
        {   // refill the ring slot with the tile PFT positions ahead
          const int nn = n + PFT;
          const double* p = (nn < NIT) ? ftile(I, nn) : ftile(I + 1, nn - NIT);
          ring[n % PFT] = p ? ldnc_v2(p) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) dmma_cc(acc[q][ct], na, *reinterpret_cast<const double2*>(&Ybuf[par][k][ct][2 * lane]));
      }
    }
  }
  __syncthreads();
  // =========================================== backward ==========================================
  {
    const long long I0 = job.sb_hi - 1;
    // window of step I0 before its update: super-block rows I0-KB+1 .. I0, right-hand side = y (in `out`)
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) {
      const long long t = tile_of((I0 + 1) * 8 - WIN, q);
#pragma unroll
      for (int ct = 0; ct < NCT; ++ct) acc[q][ct] = (t >= lo8) ? rhs_pair(job.out, job.out_rs, job.out_cs, t, ct, false) : make_double2(0.0, 0.0);
    }
    // factor tile n of step I: Ub(tile_of(8(I-KB), q), 8I + k)
    auto ftile = [&](long long I, int n) -> const double* {
      const long long t = tile_of(I * 8 - WIN, n >> 3);
      return (I >= job.sb_lo && t >= lo8) ? tile(t, I * 8 + (n & 7)) + 2 * lane : nullptr;
    };
#pragma unroll
    for (int n = 0; n < PFT; ++n) { const double* p = ftile(I0, n); ring[n] = p ? ldnc_v2(p) : make_double2(0.0, 0.0); }
    for (long long I = I0; I >= job.sb_lo; --I) {
      const int par = (int)(I & 1);
      const int rr = (int)((warp - (int)((I * 8) % 16) + 16) % 16);
      if (rr < 8) {
        const long long t = I * 8 + rr;
        const int qs = (int)((t / 16) % NSLOT);
#pragma unroll
        for (int q = 0; q < NSLOT; ++q) {
          if (q == qs) {
#pragma unroll
            for (int ct = 0; ct < NCT; ++ct) {
              store_transposed(&Ybuf[par][rr][ct][0], acc[q][ct], g, tq);
              out_pair(t, ct, acc[q][ct]);
              acc[q][ct] = (t - WIN >= lo8) ? rhs_pair(job.out, job.out_rs, job.out_cs, t - WIN, ct, false) : make_double2(0.0, 0.0);
            }
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < NIT; ++n) {
        const int q = n >> 3, k = n & 7;
        const double2 na = neg2(ring[n % PFT]);
        {
          const int nn = n + PFT;
          const double* p = (nn < NIT) ? ftile(I, nn) : ftile(I - 1, nn - NIT);
          ring[n % PFT] = p ? ldnc_v2(p) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int ct = 0; ct < NCT; ++ct) dmma_cc(acc[q][ct], na, *reinterpret_cast<const double2*>(&Ybuf[par][k][ct][2 * lane]));
      }
    }
  }
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
  const int nslot = c->kb / 2;
  const bool one = max_cols <= 8;
  const int groups = one ? 1 : (max_cols + 15) / 16;
  for (int j0 = 0; j0 < njobs; j0 += 65535) {
    WideSweepArgs b = a; b.jobs = d_jobs + j0;
    const int nj = std::min(njobs - j0, 65535);
    int rc;
    switch (nslot * 2 + (one ? 0 : 1)) {
      case 4: rc = launch_ws<2, 1, 16>(c, b, nj, groups); break;
      case 5: rc = launch_ws<2, 2, 8>(c, b, nj, groups); break;
      case 6: rc = launch_ws<3, 1, 12>(c, b, nj, groups); break;
      case 7: rc = launch_ws<3, 2, 8>(c, b, nj, groups); break;
      case 8: rc = launch_ws<4, 1, 16>(c, b, nj, groups); break;
      case 9: rc = launch_ws<4, 2, 8>(c, b, nj, groups); break;
      default: SPK_SET_ERR(c, "wide sweep: unsupported window of %d super-blocks", c->kb); return SPK_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
  }
  return SPK_OK;
}
