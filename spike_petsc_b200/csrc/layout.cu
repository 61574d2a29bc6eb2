// layout.cu -- band construction and data-movement kernels (all HBM-bound, coalesced):
//   synthetic generator, dense->tile pack, CSR(+permutation)->tile pack (the gather that applies
//   the host orderings, replacing MatPermute + MatCreateSubMatrixBanded's copy loop,
//   /root/reference/src/kspreorder.c:20 and src/matbanded.c:84-99), tile->rows unpack,
//   |a| max reduction, vector gather/scatter (VecPermute, src/kspreorder.c:122-127),
//   banded MatMult with warp-shuffle row reductions, CSR MatMult.
#include "common.cuh"

// --------------------------------------------------------------------------------------------
// Synthetic band.  One thread per row: the diagonal is delta * sum_{d!=0}|a_ij|
// accumulated for d ascending, exactly like the oracle, so CPU and GPU matrices are bit-identical.
// --------------------------------------------------------------------------------------------
__global__ void k_generate(double* __restrict__ band, BandLayout L, uint64_t seed, double delta,
                           int64_t row_offset, int64_t n_global) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t npad = L.nt * SPK_TILE;
  if (i >= npad) return;
  if (i >= L.n) { band[L.elem_off(i, i)] = 1.0; return; }
  const int k = L.k;
  const int64_t gi = row_offset + i;
  const uint64_t bw = 2 * (uint64_t)k + 1;
  double s = 0.0;
  for (int d = -k; d <= k; ++d) {
    if (d == 0) continue;
    const int64_t gj = gi + d;
    if (gj < 0 || gj >= n_global) continue;
    const double v = 2.0 * spk_u01(seed, (uint64_t)gi * bw + (uint64_t)(d + k)) - 1.0;
    band[L.elem_off(i, i + d)] = v;
    s += fabs(v);
  }
  double dg = delta * s;
  if (dg == 0.0) dg = 1.0;
  band[L.elem_off(i, i)] = dg;
}

int spk_launch_generate(spk_ctx* c, uint64_t seed, double delta) {
  const BandLayout& L = c->L;
  SPK_CUDA(c, cudaMemsetAsync(c->band, 0, sizeof(double) * (size_t)L.elems(), c->stream));
  const int64_t npad = L.nt * SPK_TILE;
  const int64_t ng = c->opts.n_global > 0 ? c->opts.n_global : L.n;
  k_generate<<<(unsigned)((npad + 127) / 128), 128, 0, c->stream>>>(c->band, L, seed, delta, c->opts.row_offset, ng);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// dense band (ROWS or DIAGS layout, device memory) -> tile-major band
// --------------------------------------------------------------------------------------------
// Columns are kept while they exist in the GLOBAL matrix, local columns [jlo, jhi): a shard's rows keep the entries
// that reach into its neighbours' columns (they land in the halo tiles = the coupling blocks).
__global__ void k_pack_dense(const double* __restrict__ src, double* __restrict__ band, BandLayout L, int layout, int64_t jlo, int64_t jhi) {
  const int64_t bw = 2 * (int64_t)L.k + 1;
  const int64_t total = L.n * bw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i, dk;
    if (layout == SPK_LAYOUT_ROWS) { i = e / bw; dk = e - i * bw; }
    else { dk = e / L.n; i = e - dk * L.n; }
    const int64_t j = i + dk - L.k;
    if (j < jlo || j >= jhi) continue;
    band[L.elem_off(i, j)] = src[e];
  }
}
__global__ void k_pad_identity(double* __restrict__ band, BandLayout L) {
  const int64_t i = L.n + threadIdx.x;
  if (i < L.nt * SPK_TILE) band[L.elem_off(i, i)] = 1.0;
}
// rows [row0, row0+nrows) of a ROWS-layout band, src pointing at the first of them (chunked host upload)
__global__ void k_pack_rows_chunk(const double* __restrict__ src, double* __restrict__ band, BandLayout L, int64_t row0, int64_t nrows,
                                  int64_t jlo, int64_t jhi) {
  const int64_t bw = 2 * (int64_t)L.k + 1;
  const int64_t total = nrows * bw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t il = e / bw, dk = e - il * bw;
    const int64_t i = row0 + il, j = i + dk - L.k;
    if (j < jlo || j >= jhi) continue;
    band[L.elem_off(i, j)] = src[e];
  }
}
int spk_launch_pack_rows_chunk(spk_ctx* c, const double* src_dev, int64_t row0, int64_t nrows) {
  const int64_t ng = c->opts.n_global > 0 ? c->opts.n_global : c->L.n;
  k_pack_rows_chunk<<<c->sm_count * 8, 256, 0, c->stream>>>(src_dev, c->band, c->L, row0, nrows, -c->opts.row_offset, ng - c->opts.row_offset);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}
int spk_launch_pack_finish(spk_ctx* c) {   // identity on the padded rows of the last tile row
  const BandLayout& L = c->L;
  if (L.nt * SPK_TILE > L.n) { k_pad_identity<<<1, 64, 0, c->stream>>>(c->band, L); SPK_KERNEL_CHECK(c); }
  return SPK_OK;
}
int spk_launch_pack_dense(spk_ctx* c, const double* src_dev, int layout) {
  const BandLayout& L = c->L;
  SPK_CUDA(c, cudaMemsetAsync(c->band, 0, sizeof(double) * (size_t)L.elems(), c->stream));
  const int64_t ng = c->opts.n_global > 0 ? c->opts.n_global : L.n;
  k_pack_dense<<<c->sm_count * 8, 256, 0, c->stream>>>(src_dev, c->band, L, layout, -c->opts.row_offset, ng - c->opts.row_offset);
  SPK_KERNEL_CHECK(c);
  if (L.nt * SPK_TILE > L.n) { k_pad_identity<<<1, 64, 0, c->stream>>>(c->band, L); SPK_KERNEL_CHECK(c); }
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// equilibration  band(i,j) <- r_i * band(i,j) * c_j  in place (MC64-style scalings exp(u_i), exp(v_j):
// the reference computes them and throws them away, src/petsc_mat_wbm.c:56).
// One thread per entry pair of a tile: coalesced 16 B accesses, the scale vectors stay in L1/L2.
// --------------------------------------------------------------------------------------------
// cs points at the scale of local column 0; columns [clo, chi) are scaled: [0, n) on a single rank, reaching kp
// entries into the neighbours' columns (the coupling blocks kept in the halo tiles) on a sharded one.
__global__ void k_scale_band(double* __restrict__ band, BandLayout L, const double* __restrict__ rs, const double* __restrict__ cs,
                             int64_t clo, int64_t chi) {
  const int64_t npairs = L.nt * (int64_t)L.tpr * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < npairs; e += (int64_t)gridDim.x * blockDim.x) {
    const int pr = (int)(e & 31);
    const int64_t t = e >> 5;
    const int64_t I = t / L.tpr;
    const int slot = (int)(t - I * L.tpr);
    const int64_t J = I + slot - L.kt;
    const int64_t i = I * 8 + (pr >> 2), j = J * 8 + (pr & 3) * 2;
    if (i >= L.n || j + 1 < clo || j >= chi) continue;
    double2* p = reinterpret_cast<double2*>(band + t * SPK_TILE_ELEMS) + pr;
    double2 v = *p;
    const double r = rs[i];
    v.x = (j >= clo && j < chi) ? r * v.x * cs[j] : v.x;
    v.y = (j + 1 >= clo && j + 1 < chi) ? r * v.y * cs[j + 1] : v.y;
    *p = v;
  }
}
int spk_launch_scale_band(spk_ctx* c, const double* rs_dev, const double* cs_dev) {
  const int64_t clo = c->opts.rank > 0 ? -(int64_t)c->kp : 0;
  const int64_t chi = c->L.n + (c->opts.rank + 1 < c->opts.nranks ? (int64_t)c->kp : 0);
  k_scale_band<<<c->sm_count * 8, 256, 0, c->stream>>>(c->band, c->L, rs_dev, cs_dev, clo, chi);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}
__global__ void k_vec_scale(double* __restrict__ out, const double* __restrict__ in, const double* __restrict__ s, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) out[e] = in[e] * s[e];
}
int spk_launch_vec_scale(spk_ctx* c, double* out, const double* in, const double* s, int64_t n) {
  k_vec_scale<<<c->sm_count * 4, 256, 0, c->stream>>>(out, in, s, n);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// tile-major band -> ROWS layout (test/debug hook)
// (columns [jlo, jhi) exist in the GLOBAL matrix: a shard's rows keep their coupling entries, like the pack kernels)
__global__ void k_unpack_rows(const double* __restrict__ band, double* __restrict__ dst, BandLayout L, int64_t jlo, int64_t jhi) {
  const int64_t bw = 2 * (int64_t)L.k + 1;
  const int64_t total = L.n * bw;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / bw, dk = e - i * bw;
    const int64_t j = i + dk - L.k;
    dst[e] = (j < jlo || j >= jhi) ? 0.0 : band[L.elem_off(i, j)];
  }
}
int spk_launch_unpack_rows(spk_ctx* c, const double* band, double* rows_dev) {
  const int64_t ng = c->opts.n_global > 0 ? c->opts.n_global : c->L.n;
  k_unpack_rows<<<c->sm_count * 8, 256, 0, c->stream>>>(band, rows_dev, c->L, -c->opts.row_offset, ng - c->opts.row_offset);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// CSR x (rowperm, colperm) -> tile-major band, one warp per (new) row, lanes over the nnz.
//   B(i, jj) = A(rowperm[i], colperm[jj])  <=>  entry (r=rowperm[i], j) lands in column icol[j];
//   kept iff |icol[j] - i| <= k  (src/matbanded.c:91).  Values are copied bit-exactly.
// --------------------------------------------------------------------------------------------
__global__ void k_pack_csr(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
                           const int* __restrict__ rowperm, const int* __restrict__ icol, double* __restrict__ band,
                           BandLayout L) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int r = rowperm ? rowperm[i] : (int)i;
    const int beg = ia[r], end = ia[r + 1];
    for (int q = beg + lane; q < end; q += 32) {
      const int j = ja[q];
      const int jj = icol ? icol[j] : j;
      const int64_t d = (int64_t)jj - i;
      if (d > L.k || d < -L.k) continue;
      band[L.elem_off(i, jj)] = a[q];
    }
  }
}
int spk_launch_pack_csr(spk_ctx* c, const CsrDev& A, const int* rowperm_dev, const int* icolperm_dev) {
  const BandLayout& L = c->L;
  SPK_CUDA(c, cudaMemsetAsync(c->band, 0, sizeof(double) * (size_t)L.elems(), c->stream));
  k_pack_csr<<<c->sm_count * 8, 256, 0, c->stream>>>(A.n, A.ia, A.ja, A.a, rowperm_dev, icolperm_dev, c->band, L);
  SPK_KERNEL_CHECK(c);
  if (L.nt * SPK_TILE > L.n) { k_pad_identity<<<1, 64, 0, c->stream>>>(c->band, L); SPK_KERNEL_CHECK(c); }
  return SPK_OK;
}

// max |a_ij| over the band (boost threshold); non-negative doubles order like their bit patterns
__global__ void k_absmax(const double* __restrict__ band, int64_t n, unsigned long long* out) {
  double m = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs(band[e]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}
int spk_launch_absmax(spk_ctx* c, const double* band, double* out_dev) {
  SPK_CUDA(c, cudaMemsetAsync(out_dev, 0, sizeof(double), c->stream));
  k_absmax<<<c->sm_count * 8, 256, 0, c->stream>>>(band, c->L.elems(), (unsigned long long*)out_dev);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// VecPermute: inverse=0: out[i] = in[idx[i]] ; inverse=1: out[idx[i]] = in[i]
// --------------------------------------------------------------------------------------------
__global__ void k_gather(const int* __restrict__ idx, int inverse, const double* __restrict__ in, double* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!inverse) out[i] = in[idx[i]];
    else out[idx[i]] = in[i];
  }
}
int spk_launch_gather(spk_ctx* c, const int* idx_dev, int inverse, const double* in, double* out, int64_t n) {
  k_gather<<<c->sm_count * 8, 256, 0, c->stream>>>(idx_dev, inverse, in, out, n);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// --------------------------------------------------------------------------------------------
// Banded MatMult  y = A x  on the tile-major band: one warp per tile row.  Lane l reads the 16 B
// (row l/4, cols 2*(l%4)..+1) of every tile of the row (512 B coalesced per tile), multiplies by
// the matching x pair, and the 4 lanes of a row are combined with two warp shuffles.
// Algorithmic bytes: B + 2*8*N.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_band_matmult(const double* __restrict__ band, BandLayout L,
                                                      const double* __restrict__ x, double* __restrict__ y,
                                                      const double* __restrict__ haloL, const double* __restrict__ haloR) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t npad = L.nt * SPK_TILE;
  for (int64_t I = warp; I < L.nt; I += nwarps) {
    const double* row = band + I * (int64_t)L.tpr * SPK_TILE_ELEMS + 2 * lane;
    double acc = 0.0;
    const int64_t J0 = I - L.kt;
#pragma unroll 4
    for (int t = 0; t < L.tpr; ++t) {
      const int64_t J = J0 + t;
      const int64_t col = J * SPK_TILE + 2 * tq;
      double2 xv;
      if (J < 0) {          // rows owned by the left-neighbour rank: haloL holds its last 8*kt entries
        if (!haloL || col + (int64_t)L.kc * SPK_TILE < 0) continue;   // (beyond the coupling block: structurally zero)
        xv = *reinterpret_cast<const double2*>(haloL + (col + (int64_t)L.kc * SPK_TILE));
      } else if (J >= L.nt) {  // right neighbour: haloR holds its first 8*kt entries
        if (!haloR || col - L.nt * SPK_TILE >= (int64_t)L.kc * SPK_TILE) continue;
        xv = *reinterpret_cast<const double2*>(haloR + (col - L.nt * SPK_TILE));
      } else {
        xv = (col + 1 < L.n) ? *reinterpret_cast<const double2*>(x + col)
                             : make_double2(col < L.n ? x[col] : 0.0, 0.0);
      }
      const double2 a = *reinterpret_cast<const double2*>(row + (int64_t)t * SPK_TILE_ELEMS);
      acc = fma(a.x, xv.x, acc);
      acc = fma(a.y, xv.y, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const int64_t i = I * SPK_TILE + g;
    if (tq == 0 && i < L.n && i < npad) y[i] = acc;
  }
}
int spk_launch_matmult(spk_ctx* c, const double* band, const double* x, double* y) {
  const int64_t warps = c->L.nt;
  int64_t blocks = (warps + 7) / 8;
  if (blocks > (int64_t)c->sm_count * 16) blocks = (int64_t)c->sm_count * 16;
  if (blocks < 1) blocks = 1;
  const bool has_left = c->opts.rank > 0, has_right = c->opts.rank + 1 < c->opts.nranks;
  k_band_matmult<<<(unsigned)blocks, 256, 0, c->stream>>>(band, c->L, x, y, has_left ? c->haloL : nullptr,
                                                          has_right ? c->haloR : nullptr);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// CSR MatMult (operator of the Krylov solve when A is sparse): warp per row, shuffle reduction
__global__ void k_csr_mult(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
                           const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int q = ia[i] + lane; q < ia[i + 1]; q += 32) s = fma(a[q], x[ja[q]], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = s;
  }
}
int spk_launch_csr_mult(spk_ctx* c, const CsrDev& A, const double* x, double* y) {
  k_csr_mult<<<c->sm_count * 8, 256, 0, c->stream>>>(A.n, A.ia, A.ja, A.a, x, y);
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}
