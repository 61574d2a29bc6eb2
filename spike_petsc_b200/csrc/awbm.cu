// awbm.cu -- approximate weighted bipartite matching on the GPU.
//
// Replaces MatGetOrdering_AWBM (/root/reference/src/petsc_mat_awbm.c:42-225), which is serial host code in the
// reference.  Like the reference it reads the CSR arrays "as if the matrix were column-major" (:47): index c walks
// CSR rows and is called a column, the ja[] entries are called rows.  The result is BIT-IDENTICAL to the serial
// algorithm (permutations are integer outputs; parity bar = exact):
//   * the streaming phases -- MatGetRowMaxAbs (:66), weights log(amax_c/|a|) (:73-80), row duals u = min (:82-87),
//     column duals v = min (w - u) (:89-94), tightness of every edge (w - u - v <= sqrt(eps), :103) and the scalings
//     exp(v)/amax, exp(u) (:212-215) -- are order independent: one warp per CSR row, the minimum over a row of the
//     OTHER index as an atomicMin on the bit pattern (weights are >= 0, so the unsigned order is the numeric order).
//     (log and exp are CUDA's, <= 1 ulp from libm: tight edges have w - u - v = 0 up to rounding while the test is
//     against 1.5e-8, so the matching does not depend on those ulps; the scalings agree to rounding.)
//   * the greedy pass (:98-112) -- "for c = 0..n-1: take the first tight row of c that no smaller column has taken" --
//     is order DEPENDENT.  It is reproduced exactly by rounds of proposals: every undecided column points at its first
//     tight row that is not finally taken and CLAIMS all its remaining ones; a row's proposal is final when the
//     proposer is the smallest undecided column that can still reach that row at all (the smallest claimer).  The
//     smallest undecided column always wins its proposal, so every round decides at least one column; in practice
//     (matrices with a heavy diagonal) nearly every column is decided in the first two rounds.  Rounds carry a stamp in
//     the upper word of the 64-bit proposal/claim keys, so nothing is reset between rounds.
//   * chains (column c waits for c-1 waits for c-2 ...) would cost one round per link: when a round decides fewer
//     than AWBM_MIN_PROGRESS columns the remaining undecided columns -- and the four repair passes of the reference
//     for columns without a free tight row (:115-193, normally a handful) -- are finished by the same serial rules on
//     the host from the downloaded state.  Decisions made on the device are final in the serial order too, so the
//     hand-over is exact at any round.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <chrono>
#include <vector>
#include "common.cuh"

#define AWBM_MIN_PROGRESS 256
#define AWBM_MAX_ROUNDS 64

__device__ __forceinline__ void atomic_min_nonneg(double* addr, double v) {   // v >= 0: unsigned order = numeric order
  atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// one warp per CSR row c: amax[c], weights, u[ja] = min weights  (:66, :73-87)
__global__ void k_awbm_weights(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
                               double* __restrict__ amax, double* __restrict__ w, double* u) {
  const int c = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (c >= n) return;
  const int lo = ia[c], hi = ia[c + 1];
  double m = 0.0;
  for (int r = lo + lane; r < hi; r += 32) m = fmax(m, fabs(a[r]));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) amax[c] = m;
  for (int r = lo + lane; r < hi; r += 32) {
    const double ar = fabs(a[r]);
    const double wt = (ar == 0.0) ? DBL_MAX : log(m / ar);
    w[r] = wt;
    atomic_min_nonneg(&u[ja[r]], wt);
  }
}

// one warp per CSR row c: v[c] = min (w - u[ja]), then the tightness of every edge of c  (:89-94, :103)
__global__ void k_awbm_duals(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ w,
                             const double* __restrict__ u, double* __restrict__ v, unsigned char* __restrict__ tight, double eps) {
  const int c = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (c >= n) return;
  const int lo = ia[c], hi = ia[c + 1];
  double m = DBL_MAX;
  for (int r = lo + lane; r < hi; r += 32) m = fmin(m, w[r] - u[ja[r]]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) v[c] = m;
  for (int r = lo + lane; r < hi; r += 32) tight[r] = ((w[r] - u[ja[r]]) - m <= eps) ? 1 : 0;
}

// keys of round `rnd`: later rounds win the atomicMin, inside a round the smaller column does
__device__ __forceinline__ unsigned long long awbm_key(int rnd, int c) {
  return ((unsigned long long)(unsigned)(AWBM_MAX_ROUNDS + 1 - rnd) << 32) | (unsigned)c;
}

// undecided column c: skip edges that are not tight or whose row is finally taken; propose to the first remaining
// row, claim every remaining one
__global__ void k_awbm_propose(int n, int rnd, const int* __restrict__ ia, const int* __restrict__ ja,
                               const unsigned char* __restrict__ tight, int* __restrict__ match, const int* __restrict__ matchR,
                               int* __restrict__ ptr, unsigned long long* prop, unsigned long long* claim) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n || match[c] != -1) return;
  const int hi = ia[c + 1];
  int p = ptr[c];
  while (p < hi && !(tight[p] && matchR[ja[p]] < 0)) ++p;
  ptr[c] = p;
  if (p == hi) { match[c] = -2; return; }   // every tight row belongs to a smaller column: left to the repair passes
  const unsigned long long key = awbm_key(rnd, c);
  atomicMin(&prop[ja[p]], key);
  for (int q = p; q < hi; ++q)
    if (tight[q] && matchR[ja[q]] < 0) atomicMin(&claim[ja[q]], key);
}

__global__ void k_awbm_accept(int n, int rnd, const int* __restrict__ ja, int* __restrict__ match, int* __restrict__ matchR,
                              const int* __restrict__ ptr, const unsigned long long* __restrict__ prop,
                              const unsigned long long* __restrict__ claim, int* __restrict__ counters) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int decided = 0, open = 0;
  if (c < n && match[c] == -1) {
    const int row = ja[ptr[c]];
    const unsigned long long key = awbm_key(rnd, c);
    if (prop[row] == key && claim[row] == key) { match[c] = row; matchR[row] = c; decided = 1; }
    else open = 1;
  }
  const unsigned md = __ballot_sync(0xffffffffu, decided), mo = __ballot_sync(0xffffffffu, open);
  if ((threadIdx.x & 31) == 0) {
    if (md) atomicAdd(&counters[0], __popc(md));
    if (mo) atomicAdd(&counters[1], __popc(mo));
  }
}

__global__ void k_awbm_scalings(int n, const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ amax,
                                double* __restrict__ sr, double* __restrict__ sc) {   // :212-215
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  sr[c] = exp(v[c]) / amax[c];
  sc[c] = exp(u[c]);
}

namespace {
struct Dev {   // device temporaries, freed on every exit path
  std::vector<void*> ptrs;
  ~Dev() { for (void* p : ptrs) cudaFree(p); }
  template <class T> cudaError_t alloc(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * (count ? count : 1));
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = (T*)p;
    return e;
  }
};
}  // namespace

// permR[match[c]] = c (:201); match, scalR, scalC, stats optional.  All pointers are host memory (orderings are setup
// calls on host CSR arrays, like spk_set_band_csr).
// stats[0] rounds on the device, [1] columns matched there, [2] columns finished by the serial greedy rule on the host,
// [3] columns that needed a repair pass (:115-193).
extern "C" int spk_awbm_csr(spk_ctx* c, int n, const int* ia, const int* ja, const double* a, int* permR, int* match_out,
                            double* scalR, double* scalC, int* stats) {
  if (!c || !ia || !ja || !a || !permR || n <= 0) return SPK_ERR_ARG;
  const long long nnz = ia[n];
  if (ia[0] != 0 || nnz < 0) { SPK_SET_ERR(c, "awbm: ia[0] = %d, ia[n] = %lld", ia[0], nnz); return SPK_ERR_ARG; }
  for (long long q = 0; q < nnz; ++q) if (ja[q] < 0 || ja[q] >= n) { SPK_SET_ERR(c, "awbm: column index %d out of range at %lld", ja[q], q); return SPK_ERR_ARG; }
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  cudaStream_t st = c->stream;
  const bool timing = getenv("SPIKE_AWBM_TIMING") != nullptr;   // debug: host wall time per phase on stderr
  auto tprev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[awbm] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tprev).count());
    tprev = now;
  };
  Dev D;
  int *d_ia, *d_ja, *d_match, *d_matchR, *d_ptr, *d_cnt;
  double *d_a, *d_amax, *d_w, *d_u, *d_v, *d_sr, *d_sc;
  unsigned char* d_tight;
  unsigned long long *d_prop, *d_claim;
  SPK_CUDA(c, D.alloc(&d_ia, (size_t)n + 1)); SPK_CUDA(c, D.alloc(&d_ja, (size_t)nnz)); SPK_CUDA(c, D.alloc(&d_a, (size_t)nnz));
  SPK_CUDA(c, D.alloc(&d_amax, (size_t)n)); SPK_CUDA(c, D.alloc(&d_w, (size_t)nnz)); SPK_CUDA(c, D.alloc(&d_u, (size_t)n));
  SPK_CUDA(c, D.alloc(&d_v, (size_t)n)); SPK_CUDA(c, D.alloc(&d_tight, (size_t)nnz));
  SPK_CUDA(c, D.alloc(&d_match, (size_t)n)); SPK_CUDA(c, D.alloc(&d_matchR, (size_t)n)); SPK_CUDA(c, D.alloc(&d_ptr, (size_t)n));
  SPK_CUDA(c, D.alloc(&d_prop, (size_t)n)); SPK_CUDA(c, D.alloc(&d_claim, (size_t)n)); SPK_CUDA(c, D.alloc(&d_cnt, 2));
  SPK_CUDA(c, cudaMemcpyAsync(d_ia, ia, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
  SPK_CUDA(c, cudaMemcpyAsync(d_ja, ja, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  SPK_CUDA(c, cudaMemcpyAsync(d_a, a, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  {
    // u = DBL_MAX (:82), match = matchR = -1 (:71, :96), ptr = ia, keys = all ones
    std::vector<double> big((size_t)n, DBL_MAX);
    SPK_CUDA(c, cudaMemcpyAsync(d_u, big.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    SPK_CUDA(c, cudaMemsetAsync(d_match, 0xff, sizeof(int) * (size_t)n, st));
    SPK_CUDA(c, cudaMemsetAsync(d_matchR, 0xff, sizeof(int) * (size_t)n, st));
    SPK_CUDA(c, cudaMemsetAsync(d_prop, 0xff, sizeof(unsigned long long) * (size_t)n, st));
    SPK_CUDA(c, cudaMemsetAsync(d_claim, 0xff, sizeof(unsigned long long) * (size_t)n, st));
    SPK_CUDA(c, cudaMemcpyAsync(d_ptr, d_ia, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    SPK_CUDA(c, cudaStreamSynchronize(st));   // `big` leaves scope
  }
  lap("alloc + upload + init");
  const double eps = sqrt(DBL_EPSILON);   // PETSC_SQRT_MACHINE_EPSILON (:59)
  const int TB = 256;
  const unsigned warp_grid = (unsigned)(((size_t)n * 32 + TB - 1) / TB), col_grid = (unsigned)((n + TB - 1) / TB);
  k_awbm_weights<<<warp_grid, TB, 0, st>>>(n, d_ia, d_ja, d_a, d_amax, d_w, d_u);
  SPK_KERNEL_CHECK(c);
  k_awbm_duals<<<warp_grid, TB, 0, st>>>(n, d_ia, d_ja, d_w, d_u, d_v, d_tight, eps);
  SPK_KERNEL_CHECK(c);
  lap("weights + duals kernels");
  int rounds = 0, on_device = 0, open = n;
  while (open > 0 && rounds < AWBM_MAX_ROUNDS) {
    ++rounds;
    SPK_CUDA(c, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), st));
    k_awbm_propose<<<col_grid, TB, 0, st>>>(n, rounds, d_ia, d_ja, d_tight, d_match, d_matchR, d_ptr, d_prop, d_claim);
    SPK_KERNEL_CHECK(c);
    k_awbm_accept<<<col_grid, TB, 0, st>>>(n, rounds, d_ja, d_match, d_matchR, d_ptr, d_prop, d_claim, d_cnt);
    SPK_KERNEL_CHECK(c);
    int cnt[2];
    SPK_CUDA(c, cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    SPK_CUDA(c, cudaStreamSynchronize(st));
    on_device += cnt[0];
    open = cnt[1];
    if (cnt[0] < AWBM_MIN_PROGRESS) break;   // a chain: the serial rule on the host is faster than a round per link
  }
  lap("matching rounds");
  if (scalR || scalC) {
    SPK_CUDA(c, D.alloc(&d_sr, (size_t)n)); SPK_CUDA(c, D.alloc(&d_sc, (size_t)n));
    k_awbm_scalings<<<col_grid, TB, 0, st>>>(n, d_u, d_v, d_amax, d_sr, d_sc);
    SPK_KERNEL_CHECK(c);
    if (scalR) SPK_CUDA(c, cudaMemcpyAsync(scalR, d_sr, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (scalC) SPK_CUDA(c, cudaMemcpyAsync(scalC, d_sc, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  std::vector<int> match((size_t)n), matchR((size_t)n);
  SPK_CUDA(c, cudaMemcpyAsync(match.data(), d_match, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  SPK_CUDA(c, cudaMemcpyAsync(matchR.data(), d_matchR, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  SPK_CUDA(c, cudaStreamSynchronize(st));

  lap("scalings + download");
  // ---- host: what is left of the greedy pass, then the reference's repair passes, all by the serial rules ----
  int host_greedy = 0, repaired = 0;
  bool need_tight = false;
  for (int col = 0; col < n && !need_tight; ++col) need_tight = match[col] < 0;
  std::vector<unsigned char> tight;
  if (need_tight) {
    tight.resize((size_t)nnz);
    SPK_CUDA(c, cudaMemcpyAsync(tight.data(), d_tight, (size_t)nnz, cudaMemcpyDeviceToHost, st));
    SPK_CUDA(c, cudaStreamSynchronize(st));
    // greedy (:98-112) for the columns the device left open, in index order (device decisions are final in that order too)
    for (int col = 0; col < n; ++col) {
      if (match[col] != -1) continue;
      for (int r = ia[col]; r < ia[col + 1]; ++r)
        if (tight[r] && matchR[ja[r]] < 0) { match[col] = ja[r]; matchR[ja[r]] = col; ++host_greedy; break; }
    }
    for (int col = 0; col < n; ++col) if (match[col] < 0) { match[col] = -1; ++repaired; }
    // one-level augmentation over tight edges (:115-140)
    for (int col = 0; col < n; ++col) {
      if (match[col] >= 0) continue;
      for (int r = ia[col]; r < ia[col + 1]; ++r) {
        if (!tight[r]) continue;
        const int c1 = matchR[ja[r]];
        if (c1 < 0) continue;   // (cannot happen after the greedy pass; the reference would index ia[-1])
        for (int r1 = ia[c1]; r1 < ia[c1 + 1]; ++r1)
          if (matchR[ja[r1]] < 0 && tight[r1]) {
            match[col] = ja[r]; matchR[ja[r]] = col; match[c1] = ja[r1]; matchR[ja[r1]] = c1;
            break;
          }
        if (match[col] >= 0) break;
      }
    }
    // non-optimal rows (:143-153)
    for (int col = 0; col < n; ++col) {
      if (match[col] >= 0) continue;
      for (int r = ia[col]; r < ia[col + 1]; ++r)
        if (matchR[ja[r]] < 0) { match[col] = ja[r]; matchR[ja[r]] = col; break; }
    }
    // non-optimal one-level augmentation (:156-178)
    for (int col = 0; col < n; ++col) {
      if (match[col] >= 0) continue;
      for (int r = ia[col]; r < ia[col + 1]; ++r) {
        const int c1 = matchR[ja[r]];
        if (c1 < 0) continue;
        for (int r1 = ia[c1]; r1 < ia[c1 + 1]; ++r1)
          if (matchR[ja[r1]] < 0) {
            match[col] = ja[r]; matchR[ja[r]] = col; match[c1] = ja[r1]; matchR[ja[r1]] = c1;
            break;
          }
        if (match[col] >= 0) break;
      }
    }
    // completion (:181-193; the row cursor persists across columns)
    for (int col = 0, r = 0; col < n; ++col) {
      if (match[col] >= 0) continue;
      for (; r < n; ++r)
        if (matchR[r] < 0) { match[col] = r; matchR[r] = col; break; }
    }
  }
  for (int col = 0; col < n; ++col)
    if (match[col] < 0 || match[col] >= n) { SPK_SET_ERR(c, "awbm: column %d unmatched", col); return SPK_ERR_STATE; }   // :196-199
  for (int col = 0; col < n; ++col) permR[match[col]] = col;   // :201
  if (match_out) std::copy(match.begin(), match.end(), match_out);
  if (stats) { stats[0] = rounds; stats[1] = on_device; stats[2] = host_greedy; stats[3] = repaired; }
  lap("host finish + permutation");
  return SPK_OK;
}
