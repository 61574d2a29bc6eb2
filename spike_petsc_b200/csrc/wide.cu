// wide.cu -- orchestration of the wide-band SPIKE path (K = 129..512): allocation, the factor stages
// (reversed tip windows -> W^(t), band LU, V^(b), reduced blocks) and the apply stages (partition sweeps,
// window corrections) in terms of the two heavy kernels k_wide_lu (wide_lu.cu) and k_wide_sweep (wide_sweep.cu).
//
// Everything dense and kp x kp (kp = 64*KB) is expressed through those two kernels:
//   V_i^(b) = last kp rows of A_i^-1 [0; B_i]   : a sweep over the last KB super-block rows of the factored
//             partition with the kp columns of B_i as right-hand sides;
//   W_i^(t) = first kp rows of A_i^-1 [C_i; 0]  : the top tip window (tipT tile rows) is copied row/column
//             reversed into its own small band (wband), factored by the same LU kernel, and swept like V;
//             the reversal of the rows is a negative stride of the sweep's input/output;
//   R_i = (I - W_{i+1}^(t) V_i^(b))^-1          : the kp x kp matrix is written in band format (rband: a dense
//             matrix of order 64*KB is a band of KB super-blocks), factored by the LU kernel, and inverted by a
//             sweep with the identity as right-hand side.
// V, W, R end up as dense row-major kp x kp blocks -- exactly what the reduced solve / boundary exchange of
// the narrow path (solve.cu, peer.cu, capi.cu) already consume.
#include "wide.cuh"
#include <algorithm>
#include <vector>
#include <cstring>

#define WIDE_CUDA(ctx, call) SPK_CUDA(ctx, call)
enum { WJ_WT = 0, WJ_VB, WJ_RED, WJ_MAIN, WJ_CORR, WJ_SITES };   // call sites of the sweep kernel (run_jobs)

int spk_wide_alloc(spk_ctx* c) {
  const BandLayout& L = c->L;
  const int P = c->P, kp = c->kp;
  const size_t kk = (size_t)kp * kp;
  c->wide_flag_parts = P + 1;
  WIDE_CUDA(c, cudaMalloc(&c->wide_flags, sizeof(unsigned long long) * (size_t)c->wide_flag_parts * WIDE_FLAGS_PER_PART));
  WIDE_CUDA(c, cudaMalloc(&c->wide_abort, sizeof(unsigned int)));
  WIDE_CUDA(c, cudaMemsetAsync(c->wide_abort, 0, sizeof(unsigned int), c->stream));
  WIDE_CUDA(c, cudaMalloc(&c->wide_zero, 4096));
  WIDE_CUDA(c, cudaMemsetAsync(c->wide_zero, 0, 4096, c->stream));
  const size_t tile_row = (size_t)L.tpr * SPK_TILE_ELEMS;
  WIDE_CUDA(c, cudaMalloc(&c->wband, sizeof(double) * tile_row * (size_t)c->tipT * P));
  WIDE_CUDA(c, cudaMalloc(&c->rband, sizeof(double) * tile_row * (size_t)(8 * c->kb) * P));
  WIDE_CUDA(c, cudaMemsetAsync(c->rband, 0, sizeof(double) * tile_row * (size_t)(8 * c->kb) * P, c->stream));
  WIDE_CUDA(c, cudaMalloc(&c->VbT, sizeof(double) * kk * P));
  std::vector<int64_t> wp(P + 1), rp(P + 1);
  for (int p = 0; p <= P; ++p) { wp[p] = (int64_t)p * c->tipT; rp[p] = (int64_t)p * 8 * c->kb; }
  WIDE_CUDA(c, cudaMalloc(&c->d_wpstart, sizeof(int64_t) * (P + 1)));
  WIDE_CUDA(c, cudaMalloc(&c->d_rpstart, sizeof(int64_t) * (P + 1)));
  WIDE_CUDA(c, cudaMemcpy(c->d_wpstart, wp.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice));
  WIDE_CUDA(c, cudaMemcpy(c->d_rpstart, rp.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice));
  c->wjobs_cap = 2 * P + 8;
  WIDE_CUDA(c, cudaMalloc(&c->d_wjobs, sizeof(WideSweepJob) * (size_t)c->wjobs_cap * WJ_SITES));
  c->h_wjobs = new std::vector<WideSweepJob>[WJ_SITES];
  return SPK_OK;
}

void spk_wide_free(spk_ctx* c) {
  auto F = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
  F(c->wide_flags); F(c->wide_abort); F(c->wide_zero); F(c->wband); F(c->rband); F(c->VbT); F(c->d_wpstart); F(c->d_rpstart);
  if (c->d_wjobs) { cudaFree(c->d_wjobs); c->d_wjobs = nullptr; }
  if (c->h_wjobs) { delete[] (std::vector<WideSweepJob>*)c->h_wjobs; c->h_wjobs = nullptr; }
  F(c->redw); c->redw_cols = 0;
  c->wide = 0; c->kb = 0;
}

// SPK_ERR_STATE when a bounded dataflow wait of a wide kernel expired (synchronises the stream)
int spk_wide_check(spk_ctx* c) {
  if (!c->wide || !c->wide_abort) return SPK_OK;
  unsigned int w = 0;
  SPK_CUDA(c, cudaMemcpyAsync(&w, c->wide_abort, sizeof(w), cudaMemcpyDeviceToHost, c->stream));
  SPK_CUDA(c, cudaStreamSynchronize(c->stream));
  if (w) { SPK_SET_ERR(c, "wide-band LU: a dataflow wait expired (abort word %u)", w); return SPK_ERR_STATE; }
  return SPK_OK;
}

// Sweep jobs of one call site (W tips, V tips, reduced inverses, partition sweeps, corrections).  The job list of a
// site is the same from call to call as long as the caller's pointers are (repeated factorisations, Krylov
// iterations): it is uploaded only when it differs from the copy already on the device, so the steady state has no
// host-to-device copy (and none of its implicit synchronisation) in front of the sweep kernels.
static int run_jobs(spk_ctx* c, const std::vector<WideSweepJob>& jobs, int max_cols, int site) {
  if (jobs.empty()) return SPK_OK;
  if ((int)jobs.size() > c->wjobs_cap) { SPK_SET_ERR(c, "wide sweep: %zu jobs exceed the job array (%d)", jobs.size(), c->wjobs_cap); return SPK_ERR_STATE; }
  WideSweepJob* dev = (WideSweepJob*)c->d_wjobs + (size_t)site * c->wjobs_cap;
  std::vector<WideSweepJob>* all = (std::vector<WideSweepJob>*)c->h_wjobs;
  std::vector<WideSweepJob>& mine = all[site];
  if (mine.size() != jobs.size() || memcmp(mine.data(), jobs.data(), sizeof(WideSweepJob) * jobs.size()) != 0) {
    SPK_CUDA(c, cudaMemcpyAsync(dev, jobs.data(), sizeof(WideSweepJob) * jobs.size(), cudaMemcpyHostToDevice, c->stream));
    mine = jobs;
  }
  return spk_wide_sweep(c, dev, (int)jobs.size(), max_cols, site == WJ_MAIN || site == WJ_CORR);
}

// --------------------------------------------------------------------------------------------
// coupling blocks -> dense row-major kp x kp:  which = 0: B_p -> out[p] (rows = last kp rows of partition p,
// columns = first kp columns to its right);  which = 1: C_p (rows = first kp rows, columns = the kp to its left)
__global__ void k_wide_extract(const double* __restrict__ band, BandLayout L, const int64_t* __restrict__ pstart, int first_part,
                               int which, double* __restrict__ out) {
  const int p = blockIdx.y + first_part;
  const int kp = L.kc * 8;
  const int64_t r0 = which == 0 ? pstart[p + 1] * 8 - kp : pstart[p] * 8;
  const int64_t c0 = which == 0 ? pstart[p + 1] * 8 : pstart[p] * 8 - kp;
  double* o = out + (size_t)p * kp * kp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kp * kp; e += gridDim.x * blockDim.x) {
    const int r = e / kp, cc = e - r * kp;
    const int dt = (int)(((c0 + cc) >> 3) - ((r0 + r) >> 3));
    o[e] = (dt <= L.kt && dt >= -L.kt) ? band[L.elem_off(r0 + r, c0 + cc)] : 0.0;
  }
}

// reversed copy of the top window of partition p: wband tile (I', J') = flipped tile (t0+W-1-I', t0+W-1-J')
__global__ void k_wide_reverse_window(const double* __restrict__ band, double* __restrict__ wband, BandLayout L,
                                      const int64_t* __restrict__ pstart, int first_part, int W) {
  const int p = blockIdx.y + first_part;
  const int64_t t0 = pstart[p];
  const int64_t ntile = (int64_t)W * L.tpr;
  const int lane = threadIdx.x & 31;
  for (int64_t tix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tix < ntile; tix += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int Ip = (int)(tix / L.tpr), slot = (int)(tix - (int64_t)Ip * L.tpr);
    const int Jp = Ip + slot - L.kt;
    double2 v = make_double2(0.0, 0.0);
    if (Jp >= 0 && Jp < W) {
      const int64_t Is = t0 + W - 1 - Ip, Js = t0 + W - 1 - Jp;
      const double2 s = *reinterpret_cast<const double2*>(band + L.tile_off(Is, Js) + 62 - 2 * lane);
      v = make_double2(s.y, s.x);
    }
    *reinterpret_cast<double2*>(wband + (((int64_t)p * W + Ip) * L.tpr + slot) * SPK_TILE_ELEMS + 2 * lane) = v;
  }
}

// out[p] = in[p]^T (dense kp x kp, row-major)
__global__ void k_wide_transpose(const double* __restrict__ in, double* __restrict__ out, int kp, int first) {
  __shared__ double t[32][33];
  const size_t base = (size_t)(blockIdx.z + first) * kp * kp;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = in[base + (size_t)(by + r) * kp + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) out[base + (size_t)(bx + r) * kp + by + threadIdx.x] = t[threadIdx.x][r];
}

// reduced matrix of interface i (= blockIdx.z + first): M = I - W V, W = Wt[i+1] (or remoteWt for the boundary
// interface), V given through VT[i] = V^T; written in band format into partition i of rband (tile rows i*8kb ..).
// One CTA per 64 x 64 block of M, warp r = tile row r of the block, operands straight from L2 as DMMA fragments.
__global__ void __launch_bounds__(256) k_wide_reduced_matrix(const double* __restrict__ Wt, const double* __restrict__ remoteWt,
                                                             const double* __restrict__ VT, double* __restrict__ rband, BandLayout L,
                                                             int first, int remote_iface) {
  const int i = blockIdx.z + first;
  const int kp = L.kc * 8, nk = L.kc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const double* W = (i == remote_iface) ? remoteWt : Wt + (size_t)(i + 1) * kp * kp;
  const double* V = VT + (size_t)i * kp * kp;
  const int rt = blockIdx.y * 8 + warp;            // tile row of M
  const int ct0 = blockIdx.x * 8;                  // first tile column
  double2 acc[8];
#pragma unroll
  for (int cidx = 0; cidx < 8; ++cidx) acc[cidx] = make_double2(0.0, 0.0);
  const double* wrow = W + (size_t)(8 * rt + g) * kp + 2 * tq;
  const double* vrow = V + (size_t)(8 * ct0 + g) * kp + 2 * tq;
  for (int k = 0; k < nk; ++k) {
    const double2 a = *reinterpret_cast<const double2*>(wrow + 8 * k);
    double2 b[8];
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) b[cidx] = *reinterpret_cast<const double2*>(vrow + (size_t)(8 * cidx) * kp + 8 * k);
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) dmma884(acc[cidx].x, acc[cidx].y, a.x, b[cidx].x);
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) dmma884(acc[cidx].x, acc[cidx].y, a.y, b[cidx].y);
  }
  const int64_t I = (int64_t)i * nk + rt;
#pragma unroll
  for (int cidx = 0; cidx < 8; ++cidx) {
    const int ct = ct0 + cidx;
    const double dx = (ct == rt && g == 2 * tq) ? 1.0 : 0.0, dy = (ct == rt && g == 2 * tq + 1) ? 1.0 : 0.0;
    *reinterpret_cast<double2*>(rband + L.tile_off(I, (int64_t)i * nk + ct) + 2 * lane) = make_double2(dx - acc[cidx].x, dy - acc[cidx].y);
  }
}

// window corrections, step 1: the scratch columns get the coupling right-hand sides at the window ends and zeros
// in between; step 3 (after the sweeps): x -= w over the windows.  job = 2*p + side as in solve.cu.
struct WideCorrArgs {
  const int64_t* pstart; int P; int W; int kp;
  const double* rtop; const double* rbot; size_t tip_stride;   // column r: rtop + r*tip_stride + p*kp
  double* w; int64_t ldw; double* x; int64_t ldx; int64_t n;
  int has_left, has_right;
};
__device__ __forceinline__ bool wide_corr_window(const WideCorrArgs& a, int p, int side, int64_t& lo, int64_t& hi, bool& use_top, bool& use_bot) {
  const int64_t t0 = a.pstart[p], t1 = a.pstart[p + 1], plen = t1 - t0;
  const bool top_on = (p > 0) || a.has_left, bot_on = (p < a.P - 1) || a.has_right;
  const bool full = 2 * (int64_t)a.W > plen;
  if (full) {
    if (side == 1 || (!top_on && !bot_on)) return false;
    lo = t0; hi = t1; use_top = top_on; use_bot = bot_on; return true;
  }
  if (side == 0) { if (!top_on) return false; lo = t0; hi = t0 + a.W; use_top = true; use_bot = false; return true; }
  if (!bot_on) return false;
  lo = t1 - a.W; hi = t1; use_top = false; use_bot = true; return true;
}
__global__ void k_wide_corr_prep(const WideCorrArgs a) {
  const int p = blockIdx.x >> 1, side = blockIdx.x & 1, col = blockIdx.y;
  int64_t lo, hi; bool ut, ub;
  if (!wide_corr_window(a, p, side, lo, hi, ut, ub)) return;
  const int64_t t0 = a.pstart[p], t1 = a.pstart[p + 1];
  const double* rt = a.rtop + (size_t)col * a.tip_stride + (size_t)p * a.kp;
  const double* rb = a.rbot + (size_t)col * a.tip_stride + (size_t)p * a.kp;
  double* w = a.w + (size_t)col * a.ldw;
  for (int64_t e = lo * 8 + threadIdx.x; e < hi * 8; e += blockDim.x) {
    double v = 0.0;
    if (ut && e < t0 * 8 + a.kp) v += rt[e - t0 * 8];
    if (ub && e >= t1 * 8 - a.kp) v += rb[e - (t1 * 8 - a.kp)];
    w[e] = v;
  }
}
__global__ void k_wide_corr_apply(const WideCorrArgs a) {
  const int p = blockIdx.x >> 1, side = blockIdx.x & 1, col = blockIdx.y;
  int64_t lo, hi; bool ut, ub;
  if (!wide_corr_window(a, p, side, lo, hi, ut, ub)) return;
  const double* w = a.w + (size_t)col * a.ldw;
  double* x = a.x + (size_t)col * a.ldx;
  for (int64_t e = lo * 8 + threadIdx.x; e < hi * 8 && e < a.n; e += blockDim.x) x[e] -= w[e];
}

// --------------------------------------------------------------------------------------------
// factor stages (called from the spk_launch_* entry points of the narrow path when c->wide)
// --------------------------------------------------------------------------------------------
// stage 0: reversed tip windows of partitions first..P-1, factored (source: the unfactored band)
int spk_wide_ul_windows(spk_ctx* c) {
  const int first = (c->opts.rank > 0) ? 0 : 1;
  const int cnt = c->P - first;
  if (cnt <= 0) return SPK_OK;
  const double* src = spk_lu_source(c);
  k_wide_reverse_window<<<dim3(64, cnt), 256, 0, c->stream>>>(src, c->wband, c->L, c->d_pstart, first, c->tipT);
  SPK_KERNEL_CHECK(c);
  return spk_wide_lu(c, c->wband, c->d_wpstart + first, cnt);
}

// stage 1: the band LU, in place on c->band (the kept original is copied over it first)
int spk_wide_band_lu(spk_ctx* c) {
  const double* src = spk_lu_source(c);
  if (src != c->band) SPK_CUDA(c, cudaMemcpyAsync(c->band, src, sizeof(double) * (size_t)c->L.elems(), cudaMemcpyDeviceToDevice, c->stream));
  return spk_wide_lu(c, c->band, c->d_pstart, c->P);
}

static int wide_wt(spk_ctx* c, int first, int cnt) {   // W^(t) of partitions first .. first+cnt-1
  if (cnt <= 0) return SPK_OK;
  const int kp = c->kp;
  k_wide_extract<<<dim3(32, cnt), 256, 0, c->stream>>>(spk_lu_source(c), c->L, c->d_pstart, first, 1, c->Wt);
  SPK_KERNEL_CHECK(c);
  std::vector<WideSweepJob> jobs;
  const long long Wsb = c->tipT / 8;
  for (int p = first; p < first + cnt; ++p) {
    WideSweepJob j{};
    j.band = c->wband; j.sb_hi = (long long)(p + 1) * Wsb; j.sb_lo = j.sb_hi - c->kb; j.sb_fwd = j.sb_lo;
    j.in = c->Wt + (size_t)p * kp * kp + (size_t)(kp - 1) * kp; j.in_rs = -kp; j.in_cs = 1;
    j.out = const_cast<double*>(j.in); j.out_rs = -kp; j.out_cs = 1;
    j.row0 = j.sb_hi * 64 - kp; j.nrow_valid = kp; j.ncols = kp;
    jobs.push_back(j);
  }
  return run_jobs(c, jobs, kp, WJ_WT);
}

static int wide_vb(spk_ctx* c, int cnt) {   // V^(b) of partitions 0 .. cnt-1 (+ their transposes)
  if (cnt <= 0) return SPK_OK;
  const int kp = c->kp;
  k_wide_extract<<<dim3(32, cnt), 256, 0, c->stream>>>(c->band, c->L, c->d_pstart, 0, 0, c->Vb);
  SPK_KERNEL_CHECK(c);
  std::vector<WideSweepJob> jobs;
  for (int p = 0; p < cnt; ++p) {
    WideSweepJob j{};
    j.band = c->band; j.sb_hi = c->h_pstart[p + 1] / 8; j.sb_lo = j.sb_hi - c->kb; j.sb_fwd = j.sb_lo;
    j.in = c->Vb + (size_t)p * kp * kp; j.in_rs = kp; j.in_cs = 1;
    j.out = const_cast<double*>(j.in); j.out_rs = kp; j.out_cs = 1;
    j.row0 = j.sb_hi * 64 - kp; j.nrow_valid = kp; j.ncols = kp;
    jobs.push_back(j);
  }
  int rc = run_jobs(c, jobs, kp, WJ_VB);
  if (rc) return rc;
  k_wide_transpose<<<dim3(kp / 32, kp / 32, cnt), dim3(32, 8), 0, c->stream>>>(c->Vb, c->VbT, kp, 0);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

static int wide_reduced(spk_ctx* c, int first, int cnt, int remote_iface) {   // R of interfaces first .. first+cnt-1
  if (cnt <= 0) return SPK_OK;
  const int kp = c->kp;
  k_wide_reduced_matrix<<<dim3(c->kb, c->kb, cnt), 256, 0, c->stream>>>(c->Wt, c->remoteWt, c->VbT, c->rband, c->L, first, remote_iface);
  SPK_KERNEL_CHECK(c);
  int rc = spk_wide_lu(c, c->rband, c->d_rpstart + first, cnt);
  if (rc) return rc;
  std::vector<WideSweepJob> jobs;
  for (int i = first; i < first + cnt; ++i) {
    WideSweepJob j{};
    j.band = c->rband; j.sb_lo = (long long)i * c->kb; j.sb_hi = j.sb_lo + c->kb; j.sb_fwd = j.sb_lo;
    j.in = nullptr; j.in_rs = 0; j.in_cs = 0;
    j.out = c->Red + (size_t)i * kp * kp; j.out_rs = kp; j.out_cs = 1;
    j.row0 = (long long)i * kp; j.nrow_valid = kp; j.ncols = kp;
    jobs.push_back(j);
  }
  return run_jobs(c, jobs, kp, WJ_RED);
}

// same contract as spk_launch_tips (tips.cu)
int spk_wide_tips(spk_ctx* c, int what) {
  const int P = c->P;
  const bool has_left = c->opts.rank > 0, has_right = c->opts.rank + 1 < c->opts.nranks;
  if (what == 1) return has_right ? wide_reduced(c, P - 1, 1, P - 1) : SPK_OK;
  if (what == 2) return has_left ? wide_wt(c, 0, 1) : SPK_OK;
  int rc = SPK_OK;
  const int nvb = (P - 1) + (has_right ? 1 : 0);
  if (what != 4) rc = wide_vb(c, nvb);
  if (rc) return rc;
  const int wfirst = has_left ? 0 : 1;
  if (!c->wt_done) rc = wide_wt(c, wfirst, P - wfirst);
  if (rc) return rc;
  if (what == 4) { c->wt_done = 1; return SPK_OK; }
  const bool with_bnd = (what == 3 && has_right);
  return wide_reduced(c, 0, (P - 1) + (with_bnd ? 1 : 0), with_bnd ? P - 1 : -1);
}

// --------------------------------------------------------------------------------------------
// apply stages
// --------------------------------------------------------------------------------------------
// g = D^-1 b: one job per partition, nrhs columns (column r at b + r*ld)
int spk_wide_main_sweep(spk_ctx* c, const double* b, double* x, int nrhs, int64_t ld) {
  std::vector<WideSweepJob> jobs;
  for (int p = 0; p < c->P; ++p) {
    WideSweepJob j{};
    j.band = c->band; j.sb_lo = c->h_pstart[p] / 8; j.sb_hi = c->h_pstart[p + 1] / 8; j.sb_fwd = j.sb_lo;
    j.in = b; j.in_rs = 1; j.in_cs = ld; j.out = x; j.out_rs = 1; j.out_cs = ld;
    j.row0 = 0; j.nrow_valid = c->L.n; j.ncols = nrhs;
    jobs.push_back(j);
  }
  return run_jobs(c, jobs, nrhs, WJ_MAIN);
}

// x_i -= A_i^-1 [r_top; 0] + A_i^-1 [0; r_bot] over the truncation windows; tips: column r at rtop/rbot + r*tip_stride;
// work: nrhs scratch columns of leading dimension ldw >= padded rows
int spk_wide_corrections(spk_ctx* c, double* x, int nrhs, int64_t ld, const double* rtop, const double* rbot, size_t tip_stride,
                         double* work, int64_t ldw) {
  WideCorrArgs a;
  a.pstart = c->d_pstart; a.P = c->P; a.W = c->tipT; a.kp = c->kp;
  a.rtop = rtop; a.rbot = rbot; a.tip_stride = tip_stride; a.w = work; a.ldw = ldw; a.x = x; a.ldx = ld; a.n = c->L.n;
  a.has_left = c->opts.rank > 0; a.has_right = c->opts.rank + 1 < c->opts.nranks;
  k_wide_corr_prep<<<dim3(2 * c->P, nrhs), 256, 0, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  std::vector<WideSweepJob> jobs;
  const int kt8 = 8 * c->kb;   // tile rows of a tip
  for (int p = 0; p < c->P; ++p) {
    const int64_t t0 = c->h_pstart[p], t1 = c->h_pstart[p + 1], plen = t1 - t0;
    const bool top_on = (p > 0) || a.has_left, bot_on = (p < c->P - 1) || a.has_right;
    const bool full = 2 * (int64_t)c->tipT > plen;
    for (int side = 0; side < 2; ++side) {
      int64_t lo, hi, flo;
      if (full) {
        if (side == 1 || (!top_on && !bot_on)) continue;
        lo = t0; hi = t1; flo = top_on ? t0 : t1 - kt8;
      } else if (side == 0) {
        if (!top_on) continue;
        lo = t0; hi = t0 + c->tipT; flo = lo;
      } else {
        if (!bot_on) continue;
        lo = t1 - c->tipT; hi = t1; flo = t1 - kt8;
      }
      WideSweepJob j{};
      j.band = c->band; j.sb_lo = lo / 8; j.sb_hi = hi / 8; j.sb_fwd = flo / 8;
      j.in = work; j.in_rs = 1; j.in_cs = ldw; j.out = work; j.out_rs = 1; j.out_cs = ldw;
      j.row0 = 0; j.nrow_valid = ldw; j.ncols = nrhs;
      jobs.push_back(j);
    }
  }
  int rc = run_jobs(c, jobs, nrhs, WJ_CORR);
  if (rc) return rc;
  k_wide_corr_apply<<<dim3(2 * c->P, nrhs), 256, 0, c->stream>>>(a);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// reduced solve for wide bands, all right-hand sides at once.  Per interface i (partition i below, i+1 above):
//     t = g_t - W g_b,   x_t = R t,   x_b = g_b - V x_t,   r_bot(i) = B_i x_t,   r_top(i+1) = C_{i+1} x_b
// Each product is (kp x kp) times (kp x nrhs): k_wide_mv runs it on the FP64 tensor cores, one CTA per 64 rows of
// one interface, the matrix tiles straight from L2 as left fragments (dense row-major W, R, V or the coupling
// blocks where they lie in the band), the vectors as right fragments from a column-major scratch (ld = kp).
// The narrow path's k_reduced_solve (solve.cu) does the same with mat-vecs; at kp = 512 and 32 columns that would
// re-read the three 2 MB matrices per column.
struct WideMvArgs {
  int mode;                           // 0: dense row-major (ld = kp)   1: B_i in the band   2: C_{i+1} in the band
  const double* M; size_t m_stride; int m_off;      // dense: matrix of interface i = M + (i + m_off) * m_stride
  const double* M_remote; int remote_iface;         // dense: interface remote_iface uses M_remote instead
  const double* band; BandLayout L; const int64_t* pstart;
  const double* X; const double* base; double* out;  // per interface blocks of kp x ncap (column-major, ld = kp) ...
  size_t v_stride;                                   // ... v_stride apart
  double* out_ext; size_t oe_iface; long long oe_ld;  // or (out == nullptr) an external output: out_ext + i*oe_iface, column c at + c*oe_ld
  int kp, nrhs, first; double alpha;                 // out = base + alpha * M X
};
__global__ void __launch_bounds__(256) k_wide_mv(const WideMvArgs a) {
  const int i = blockIdx.y + a.first;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int kp = a.kp, nk = kp >> 3;
  const int rt = blockIdx.x * 8 + warp;
  const int c0 = blockIdx.z * 32;
  const double* X = a.X + (size_t)i * a.v_stride;
  double2 acc[4];
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) acc[ct] = make_double2(0.0, 0.0);
  int klo = 0, khi = nk;
  const double* mrow = nullptr; long long mstep = 0;
  if (a.mode == 0) {
    const double* M = (i == a.remote_iface) ? a.M_remote : a.M + (size_t)(i + a.m_off) * a.m_stride;
    mrow = M + (size_t)(8 * rt + g) * kp + 2 * tq; mstep = 8;
  } else {
    const int64_t tb = a.pstart[i + 1];
    if (a.mode == 1) { khi = rt + 1; mrow = a.band + a.L.tile_off(tb - a.L.kc + rt, tb) + 2 * lane; }   // B_i: tiles k <= rt
    else { klo = rt; mrow = a.band + a.L.tile_off(tb + rt, tb - a.L.kc) + 2 * lane; }                   // C_{i+1}: tiles k >= rt
    mstep = SPK_TILE_ELEMS;
  }
  for (int k = klo; k < khi; ++k) {
    const double2 av = *reinterpret_cast<const double2*>(mrow + (long long)k * mstep);
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) {
      const int col = c0 + 8 * ct + g;
      const double2 bv = (col < a.nrhs) ? *reinterpret_cast<const double2*>(X + (size_t)col * kp + 8 * k + 2 * tq) : make_double2(0.0, 0.0);
      dmma_cc(acc[ct], av, bv);
    }
  }
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    const int col = c0 + 8 * ct + 2 * tq, row = 8 * rt + g;
    double vx = a.alpha * acc[ct].x, vy = a.alpha * acc[ct].y;
    if (a.base) {
      const double* b = a.base + (size_t)i * a.v_stride;
      if (col < a.nrhs) vx += b[(size_t)col * kp + row];
      if (col + 1 < a.nrhs) vy += b[(size_t)(col + 1) * kp + row];
    }
    if (a.out) {
      double* o = a.out + (size_t)i * a.v_stride;
      if (col < a.nrhs) o[(size_t)col * kp + row] = vx;
      if (col + 1 < a.nrhs) o[(size_t)(col + 1) * kp + row] = vy;
    } else {
      double* o = a.out_ext + (size_t)i * a.oe_iface;
      if (col < a.nrhs) o[(long long)col * a.oe_ld + row] = vx;
      if (col + 1 < a.nrhs) o[(long long)(col + 1) * a.oe_ld + row] = vy;
    }
  }
}
// g_b(i) = last kp entries of partition i, g_t(i) = first kp entries of partition i+1 (or the right rank's, remote)
__global__ void k_wide_red_gather(const double* __restrict__ x, int64_t ldx, int64_t n, const int64_t* __restrict__ pstart, int kp,
                                  int first, int remote_iface, const double* __restrict__ remote_gt, double* gb, double* gt, size_t v_stride) {
  const int i = blockIdx.x + first, col = blockIdx.y;
  const int64_t tb = pstart[i + 1] * 8;
  for (int e = threadIdx.x; e < kp; e += blockDim.x) {
    gb[(size_t)i * v_stride + (size_t)col * kp + e] = x[(size_t)col * ldx + tb - kp + e];
    double v;
    if (i == remote_iface) v = remote_gt[(size_t)col * kp + e];
    else v = (tb + e < n) ? x[(size_t)col * ldx + tb + e] : 0.0;
    gt[(size_t)i * v_stride + (size_t)col * kp + e] = v;
  }
}

int spk_wide_reduced_solve(spk_ctx* c, const double* x, int nrhs, int64_t ld, double* rtop, double* rbot, size_t tip_stride) {
  const bool has_right = c->opts.rank + 1 < c->opts.nranks;
  const int nif = (c->P - 1) + (has_right ? 1 : 0);
  if (nif <= 0) return SPK_OK;
  const int kp = c->kp, P = c->P;
  const int bnd = has_right ? P - 1 : -1;
  if (c->redw_cols < nrhs) {
    if (c->redw) { cudaFree(c->redw); c->redw = nullptr; }
    c->redw_cols = 0;
    SPK_CUDA(c, cudaMalloc(&c->redw, sizeof(double) * 5 * (size_t)P * kp * nrhs));
    c->redw_cols = nrhs;
  }
  const size_t vs = (size_t)kp * c->redw_cols, blk = (size_t)P * vs;
  double *gb = c->redw, *gt = gb + blk, *tv = gt + blk, *xt = tv + blk, *xb = xt + blk;
  k_wide_red_gather<<<dim3(nif, nrhs), 256, 0, c->stream>>>(x, ld, c->L.n, c->d_pstart, kp, 0, bnd, c->remoteGtop, gb, gt, vs);
  SPK_KERNEL_CHECK(c);
  WideMvArgs a{};
  a.band = c->band; a.L = c->L; a.pstart = c->d_pstart; a.kp = kp; a.nrhs = nrhs; a.first = 0; a.v_stride = vs;
  a.m_stride = (size_t)kp * kp; a.remote_iface = -1;
  const dim3 grid(kp / 64, nif, (nrhs + 31) / 32);
  // t = g_t - W g_b
  a.mode = 0; a.M = c->Wt; a.m_off = 1; a.M_remote = c->remoteWt; a.remote_iface = bnd; a.X = gb; a.base = gt; a.out = tv; a.alpha = -1.0;
  k_wide_mv<<<grid, 256, 0, c->stream>>>(a); SPK_KERNEL_CHECK(c);
  // x_t = R t
  a.M = c->Red; a.m_off = 0; a.remote_iface = -1; a.X = tv; a.base = nullptr; a.out = xt; a.alpha = 1.0;
  k_wide_mv<<<grid, 256, 0, c->stream>>>(a); SPK_KERNEL_CHECK(c);
  // x_b = g_b - V x_t
  a.M = c->Vb; a.X = xt; a.base = gb; a.out = xb; a.alpha = -1.0;
  k_wide_mv<<<grid, 256, 0, c->stream>>>(a); SPK_KERNEL_CHECK(c);
  // r_bot(i) = B_i x_t
  a.mode = 1; a.X = xt; a.base = nullptr; a.out = nullptr; a.alpha = 1.0;
  a.out_ext = rbot; a.oe_iface = (size_t)kp; a.oe_ld = (long long)tip_stride;
  k_wide_mv<<<grid, 256, 0, c->stream>>>(a); SPK_KERNEL_CHECK(c);
  // r_top(i+1) = C_{i+1} x_b for the local interfaces; the boundary one hands x_b to the right rank instead
  if (P - 1 > 0) {
    a.mode = 2; a.X = xb; a.out_ext = rtop + kp;
    k_wide_mv<<<dim3(kp / 64, P - 1, (nrhs + 31) / 32), 256, 0, c->stream>>>(a); SPK_KERNEL_CHECK(c);
  }
  if (has_right) {
    if (nrhs > c->bnd_cols) { SPK_SET_ERR(c, "sharded wide solve: %d right-hand sides exceed the boundary buffers (%d)", nrhs, c->bnd_cols); return SPK_ERR_STATE; }
    SPK_CUDA(c, cudaMemcpy2DAsync(c->xbBoundary, sizeof(double) * kp, xb + (size_t)bnd * vs, sizeof(double) * kp, sizeof(double) * kp, nrhs,
                                  cudaMemcpyDeviceToDevice, c->stream));
  }
  return SPK_OK;
}
