"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs; size-independent properties at BASELINE.json's full sizes.

Bars (north_star): band structure / permutation application bit-exact; solution vectors within
1e-10 relative of the reference CPU path (exact band LU solve); Krylov iteration counts within +-1.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
RTOL = 1e-10  # north_star: solution vectors within 1e-10 relative error (fp64)


def relerr(x, ref):
    return np.linalg.norm(x - ref) / np.linalg.norm(ref)


# ------------------------------------------------------------------ layout / generator (bit exact)
@pytest.mark.parametrize("n,k", [(64, 5), (1000, 10), (4099, 37), (2048, 100), (777, 128)])
def test_synthetic_band_bit_exact(spk, oracle, n, k):
    S = spk.Spike()
    S.set_band_synthetic(n, k, seed=20140601, delta=1.2)
    np.testing.assert_array_equal(S.get_band_rows(), oracle.gen_band(n, k, 20140601, 1.2))
    S.close()


def test_generator_golden_on_gpu(spk):
    g = GOLD["synthetic_n64_k5"]
    S = spk.Spike()
    S.set_band_synthetic(64, 5, seed=g["seed"], delta=g["delta"])
    a = S.get_band_rows()
    np.testing.assert_array_equal(a[0], g["band_row0"])
    np.testing.assert_array_equal(a[17], g["band_row17"])
    np.testing.assert_array_equal(a[63], g["band_row63"])
    S.close()


@pytest.mark.parametrize("layout", ["rows", "diags"])
def test_dense_pack_round_trip_bit_exact(spk, oracle, layout):
    n, k = 1237, 21
    a = oracle.gen_band(n, k, seed=3)
    S = spk.Spike()
    if layout == "rows":
        S.set_band_dense(a, k, spk.LAYOUT_ROWS)
    else:
        S.set_band_dense(np.ascontiguousarray(a.T), k, spk.LAYOUT_DIAGS)
    np.testing.assert_array_equal(S.get_band_rows(), a)
    S.close()


def test_host_rows_upload_is_chunked_and_bit_exact(spk, oracle):
    """spk_set_band_dense(host, ROWS) streams the band through two chunk buffers (copy stream + pack kernel);
    with 64-row chunks a 1237-row band takes 20 of them, the last one ragged."""
    import ctypes as C
    L = spk.lib()
    L.spk_debug_set_pack_chunk_rows.argtypes = [C.c_int64]
    n, k = 1237, 21
    a = oracle.gen_band(n, k, seed=3)
    try:
        L.spk_debug_set_pack_chunk_rows(64)
        S = spk.Spike()
        S.set_band_dense(a, k, spk.LAYOUT_ROWS)
        np.testing.assert_array_equal(S.get_band_rows(), a)
        S.factor()
        b = oracle.band_mult(a, np.ones(n))
        assert relerr(S.solve(b), np.ones(n)) < 1e-12
        S.close()
    finally:
        L.spk_debug_set_pack_chunk_rows(0)


# ------------------------------------------------------------------ MatMult
@pytest.mark.parametrize("n,k", [(100, 3), (5000, 10), (20000, 50), (9999, 100)])
def test_matmult_vs_oracle(spk, oracle, n, k):
    a = oracle.gen_band(n, k)
    x = oracle.gen_vec(n, 5) - 0.5
    S = spk.Spike()
    S.set_band_dense(a, k)
    y = S.mult(x)
    ref = oracle.band_mult(a, x)
    assert relerr(y, ref) < 1e-14
    S.close()


# ------------------------------------------------------------------ factor: block-LU entries vs oracle
@pytest.mark.parametrize("n,k", [(512, 10), (4096, 20), (3000, 50), (4096, 100), (2500, 128), (1001, 64)])
def test_lu_factors_match_oracle_single_partition(spk, oracle, n, k):
    """The kernel stores the block LU grouped by 8 pivots (A~ below, Ub = D^-1 A~ above, D^-1 on the
    diagonal tiles); oracle.block_lu restates exactly that on the CPU, with the same Schur
    complements, pivots and boosting rule as the scalar no-pivot LU (oracle.band_lu)."""
    a = oracle.gen_band(n, k, seed=n + k)
    wide, nb, kw = oracle.block_lu(a)
    S = spk.Spike(partitions=1, tip_tiles=-1)
    S.set_band_dense(a, k)
    S.factor()
    f = S.get_band_rows()
    ref = wide[:, kw - k:kw + k + 1]
    assert np.abs(f - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    assert S.view()["boosted_pivots"] == 0 and nb == 0
    # and the block factors solve like the scalar ones
    lu, _ = oracle.band_lu(a)
    b = oracle.band_mult(a, np.ones(n))
    assert relerr(S.solve(b), oracle.band_solve(lu, b)) < RTOL
    S.close()


# ------------------------------------------------------------------ SPIKE solve vs exact band solve
CASES = [
    # n, k, partitions, tip_tiles (-1 = full windows), delta
    (4000, 10, 4, -1, 1.2),
    (100_000, 10, 0, 0, 1.2),      # BASELINE config[0] shape: N=100k K=10
    (100_000, 10, 592, 12, 1.2),   # ... at the partitions / window bench.py runs it with (tools/side_ab.py sweep)
    (100_000, 10, 592, 12, 1.0),
    (20_000, 50, 8, -1, 1.2),
    (20_001, 37, 5, -1, 1.2),      # ragged: n not a multiple of 8, k not a multiple of 8
    (60_000, 100, 12, -1, 1.2),
    (60_000, 100, 12, 0, 1.2),     # auto truncation window
    (50_000, 128, 6, 0, 1.2),      # widest band of the register-window kernel
    (30_000, 50, 16, 0, 1.0),      # weak dominance (delta = 1.0)
]


@pytest.mark.parametrize("n,k,P,tip,delta", CASES)
def test_spike_solve_matches_reference_cpu_path(spk, oracle, n, k, P, tip, delta):
    a = oracle.gen_band(n, k, delta=delta)
    lu, _ = oracle.band_lu(a)
    S = spk.Spike(partitions=P, tip_tiles=tip)
    S.set_band_dense(a, k)
    S.factor()
    for u in (np.ones(n), oracle.gen_vec(n, 20140601)):      # u = 1 and -random_exact_sol (src/testbed2.c:112-121)
        b = oracle.band_mult(a, u)
        x = S.solve(b)
        xref = oracle.band_solve(lu, b)
        assert relerr(x, xref) < RTOL, (relerr(x, xref), S.view())
        assert relerr(x, u) < 1e-9
    S.close()


def test_multiple_rhs(spk, oracle):
    n, k = 12_000, 30
    a = oracle.gen_band(n, k)
    lu, _ = oracle.band_lu(a)
    U = np.stack([oracle.gen_vec(n, s) for s in range(4)])
    Bm = np.stack([oracle.band_mult(a, u) for u in U])
    S = spk.Spike(partitions=6)
    S.set_band_dense(a, k)
    S.factor()
    X = S.solve(Bm, nrhs=4)
    for r in range(4):
        assert relerr(X[r], oracle.band_solve(lu, Bm[r])) < RTOL
    S.close()


@pytest.mark.parametrize("n,k,P,nrhs", [(20_000, 100, 4, 32), (9_001, 37, 3, 9), (4_096, 10, 1, 2), (30_000, 50, 8, 40),
                                         (6_000, 128, 2, 17)])
def test_multi_rhs_tensor_core_sweeps(spk, oracle, n, k, P, nrhs):
    """nrhs >= 2 runs the partition sweeps for all columns at once (msweep.cu, 8 columns per warp on DMMA):
    same answers as the exact band solve, ragged column counts and row counts included."""
    a = oracle.gen_band(n, k)
    lu, _ = oracle.band_lu(a)
    U = np.stack([oracle.gen_vec(n, 100 + s) for s in range(nrhs)])
    Bm = np.stack([oracle.band_mult(a, u) for u in U])
    S = spk.Spike(partitions=P)
    S.set_band_dense(a, k)
    S.factor()
    X = S.solve(Bm, nrhs=nrhs)
    x1 = S.solve(Bm[nrhs - 1])                       # the single-column kernels agree with the block path
    for r in range(nrhs):
        assert relerr(X[r], oracle.band_solve(lu, Bm[r])) < RTOL, r
    assert relerr(X[nrhs - 1], x1) < 1e-13
    S.close()


def test_refactor_reads_the_kept_original(spk, oracle):
    """With spk_keep_original(ctx,1) the LU reads the kept unfactored band and writes the factors into the working band,
    so spk_factor can be repeated (PCSetUp after a reset) and gives bit-identical factors and solutions."""
    n, k = 20_000, 40
    a = oracle.gen_band(n, k)
    b = oracle.band_mult(a, np.ones(n))
    S = spk.Spike(partitions=5)
    S.keep_original(True)
    S.set_band_dense(a, k)
    S.factor()
    f1, x1 = S.get_band_rows().copy(), S.solve(b).copy()
    S.factor()
    S.factor()
    np.testing.assert_array_equal(S.get_band_rows(), f1)
    np.testing.assert_array_equal(S.solve(b), x1)
    assert relerr(x1, np.ones(n)) < 1e-12
    assert relerr(S.mult(np.ones(n)), b) < 1e-14
    S.close()


def test_gpu_tips_match_oracle_tips(spk, oracle):
    """Same partitioning on CPU and GPU -> identical truncated-SPIKE answer (not only the exact one)."""
    n, k, P = 24_000, 40, 6
    a = oracle.gen_band(n, k, delta=1.0)
    b = oracle.band_mult(a, np.ones(n))
    So = oracle.Spike(n, k, P, align=8, tip_rows=0)
    So.factor(a)
    xo = So.solve(b)
    S = spk.Spike(partitions=P, tip_tiles=-1)
    S.set_band_dense(a, k)
    S.factor()
    assert relerr(S.solve(b), xo) < 1e-12
    S.close()


def test_boosting_counts_zero_pivots(spk, oracle):
    n, k = 4096, 16
    a = oracle.gen_band(n, k)
    a[100, :] = 0.0
    a[:, :] = a  # row 100 entirely zero -> pivot 100 is exactly zero after elimination
    for d in range(1, k + 1):
        a[100 - d, k + d] = 0.0  # and nothing in column 100 above it
    S = spk.Spike(partitions=1, boost_rel=1e-10)
    S.set_band_dense(a, k)
    S.factor()
    info = S.view()
    assert info["boosted_pivots"] >= 1
    lu, nb = oracle.band_lu(a, boost=1e-10 * info["anorm_max"])
    assert nb == info["boosted_pivots"]
    S.close()


@pytest.mark.parametrize("kind", ["second_difference", "weak_rows"])
def test_pivot_block_inverse_tiers(spk, oracle, kind):
    """The LU kernel inverts 8x8 pivot blocks by Newton-Schulz on the tensor cores, started from a Jacobi
    iterate; blocks that are not diagonally dominant must fall through to the FP32 Gauss-Jordan start (and,
    for tiny pivots, to the exact FP64 Gauss-Jordan) and still reproduce the no-pivot LU of the reference
    path (oracle.band_lu / oracle.block_lu)."""
    n = 4096
    if kind == "second_difference":      # tridiag(-1, 2, -1) inside a K=9 band: |offdiag|/|diag| = 1/2 in every block
        k = 9
        a = np.zeros((n, 2 * k + 1))
        a[:, k] = 2.0
        a[1:, k - 1] = -1.0
        a[:-1, k + 1] = -1.0
    else:                                # random band, rows only weakly dominant (delta = 0.35 of the off-diagonal sum)
        k = 20
        a = oracle.gen_band(n, k, delta=0.35, seed=7)
    wide, nb, kw = oracle.block_lu(a)
    S = spk.Spike(partitions=1, tip_tiles=-1)
    S.set_band_dense(a, k)
    S.factor()
    f = S.get_band_rows()
    ref = wide[:, kw - k:kw + k + 1]
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(f - ref).max() <= 1e-9 * scale, np.abs(f - ref).max() / scale
    lu, _ = oracle.band_lu(a)
    u = oracle.gen_vec(n, 11)
    b = oracle.band_mult(a, u)
    xref = oracle.band_solve(lu, b)
    assert relerr(S.solve(b), xref) < 1e-9
    S.close()


def test_error_paths(spk, oracle):
    S = spk.Spike()
    with pytest.raises(spk.SpikeError):
        S.factor()                                   # no band yet
    a = oracle.gen_band(400, 4)
    S.set_band_dense(a, 4)
    with pytest.raises(spk.SpikeError):
        S.solve(np.ones(400))                        # solve before factor
    S.factor()
    with pytest.raises(spk.SpikeError):
        S.factor()                                   # in-place factorisation cannot be repeated
    with pytest.raises(spk.SpikeError):
        S.mult(np.ones(400))                         # band overwritten, keep_original not requested
    S2 = spk.Spike()
    with pytest.raises(spk.SpikeError):
        S2.set_band_synthetic(100_000, 1000)         # wider than 512: not supported
    S.close(); S2.close()


# ------------------------------------------------------------------ CSR (+permutation) -> band, bit exact
def _sparse_case(n, k, seed, extra=2):
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for i in range(n):
        for d in range(-k, k + 1):
            j = i + d
            if 0 <= j < n and (d == 0 or rng.uniform() < 0.6):
                rows.append(i); cols.append(j); vals.append(rng.uniform(-1, 1) if d else 2.5 * k)
    for _ in range(extra * n):
        i, j = rng.integers(0, n, 2)
        rows.append(i); cols.append(j); vals.append(0.01 * rng.standard_normal())
    A = sp.csr_matrix(sp.coo_matrix((vals, (rows, cols)), shape=(n, n)))
    A.sum_duplicates(); A.sort_indices()
    return A


@pytest.mark.parametrize("kmax,frac", [(50, 0.95), (12, 1.0), (5, 0.9), (12, 0.5)])
def test_band_selection_and_extraction_bit_exact(spk, oracle, kmax, frac):
    n = 3000
    A = _sparse_case(n, 12, seed=4)
    k_ref, f_ref = oracle.band_select(A.indptr, A.indices, A.data, kmax, frac)
    S = spk.Spike()
    k, f = S.set_band_csr(A.indptr, A.indices, A.data, kmax, frac)
    assert (k, f) == (k_ref, f_ref)                   # bit-exact k and norm fraction (src/matbanded.c:104-105)
    if k > 0:
        np.testing.assert_array_equal(S.get_band_rows(), oracle.csr_to_band(A.indptr, A.indices, A.data, k))
    S.close()


def test_permuted_band_bit_exact(spk, oracle):
    """MatPermute (src/kspreorder.c:20) fused with the band extraction, orderings from the oracle."""
    n = 2500
    A = _sparse_case(n, 9, seed=8)
    rng = np.random.default_rng(3)
    q = rng.permutation(n).astype(np.int32)
    Aq = sp.csr_matrix(A[q, :][:, q]); Aq.sort_indices()          # hide the band behind a symmetric permutation
    inv = np.argsort(q).astype(np.int32)                           # the ordering that recovers it
    ib, jb, b = oracle.mat_permute_csr(Aq.indptr, Aq.indices, Aq.data, inv, inv)
    k_ref, f_ref = oracle.band_select(ib, jb, b, 20, 0.99)
    S = spk.Spike()
    k, f = S.set_band_csr(Aq.indptr, Aq.indices, Aq.data, 20, 0.99, rowperm=inv, colperm=inv)
    assert (k, f) == (k_ref, f_ref)
    np.testing.assert_array_equal(S.get_band_rows(), oracle.csr_to_band(ib, jb, b, k))
    S.close()


def test_vec_permute_bit_exact(spk, oracle):
    rng = np.random.default_rng(0)
    n = 100_003
    idx = rng.permutation(n).astype(np.int32)
    x = rng.standard_normal(n)
    S = spk.Spike()
    y = S.permute(idx, x, inverse=False)
    np.testing.assert_array_equal(y, oracle.vec_permute(x, idx, False))
    np.testing.assert_array_equal(S.permute(idx, y, inverse=True), x)
    S.close()


def test_wbm_3x3_permutation_applied_on_gpu(spk, oracle):
    """Reference KAT (src/wbm.c:483-497): MC64's column IS applied by the GPU gather gives the
    matrix the reference's MatPermute would give."""
    g = GOLD["wbm3x3"]
    n = 3
    ib, jb, b = oracle.mat_permute_csr(g["ia"], g["ja"], g["a"], g["row_is"], g["col_is"])
    S = spk.Spike()
    S.set_operator_csr(g["ia"], g["ja"], g["a"], rowperm=g["row_is"], colperm=g["col_is"])
    P = sp.csr_matrix((b, jb, ib), shape=(n, n)).toarray()
    np.testing.assert_array_equal(P, sp.csr_matrix((g["a"], g["ja"], g["ia"]), shape=(n, n)).toarray()[:, g["col_is"]])
    S.close()


# ------------------------------------------------------------------ equilibration
@pytest.mark.parametrize("scales", ["ideal", "mc64"])
def test_equilibration_removes_spurious_boosts(spk, oracle, scales):
    """A = D1 T D2 with T diagonally dominant and row scales over sixteen decades: the boosting rule
    (|pivot| < boost_rel * max|a|) mistakes the small rows for singular ones (on the GPU only where the
    tensor-core pivot-block inverse falls back to the scalar elimination).  With the band equilibrated by
    spk_set_scaling -- by 1/D1, 1/D2, or by exp(u), exp(v) of the reference's own MC64 job 5, the vector
    src/petsc_mat_wbm.c:56 throws away -- no pivot is boosted and spk_solve returns the exact band solve of the
    ORIGINAL system."""
    if scales == "mc64" and not oracle.have_mc64():
        pytest.skip("reference MC64 not built (oracle/_ref)")
    n, k = 6000, 12
    t = oracle.gen_band(n, k)
    rng = np.random.default_rng(7)
    d1, d2 = 10.0 ** rng.uniform(-8, 8, n), 10.0 ** rng.uniform(-0.5, 0.5, n)   # (wide column scales would make x itself ill-determined)
    a = np.zeros_like(t)
    rows, cols, vals = [], [], []
    for d in range(-k, k + 1):
        lo, hi = max(0, -d), min(n, n - d)
        a[lo:hi, d + k] = d1[lo:hi] * t[lo:hi, d + k] * d2[lo + d:hi + d]
        rows.append(np.arange(lo, hi)); cols.append(np.arange(lo, hi) + d); vals.append(a[lo:hi, d + k])
    u = oracle.gen_vec(n, 5)
    b = oracle.band_mult(a, u)
    lu, _ = oracle.band_lu(a)
    xref = oracle.band_solve(lu, b)
    # the scalar rule restated by the oracle boosts hundreds of healthy pivots of the small rows ...
    assert oracle.band_lu(a, boost=1e-13 * np.abs(a).max())[1] > 100
    if scales == "ideal":
        r, c = 1.0 / d1, 1.0 / d2
    else:
        m = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        m.sort_indices()
        _, col_is, num, dw = oracle.wbm(m.indptr, m.indices, m.data)
        assert num == n and np.array_equal(col_is, np.arange(n))          # the dominant diagonal is the matching
        # the wrapper hands CSR to MC64 as CSC (src/petsc_mat_wbm.c:29,52): MC64 sees A^T, so its row scaling u
        # (dw[0:n]) scales the COLUMNS of A and its column scaling v (dw[n:2n]) the rows
        r, c = np.exp(dw[n:2 * n]), np.exp(dw[:n])
        sc = sp.diags(r) @ m @ sp.diags(c)
        assert abs(sc).max() <= 1.0 + 1e-10 and np.allclose(abs(sc.diagonal()), 1.0, rtol=1e-10)
    # ... and none once the band is equilibrated
    a_sc = np.zeros_like(a)
    for d in range(-k, k + 1):
        lo, hi = max(0, -d), min(n, n - d)
        a_sc[lo:hi, d + k] = r[lo:hi] * a[lo:hi, d + k] * c[lo + d:hi + d]
    assert oracle.band_lu(a_sc, boost=1e-13 * np.abs(a_sc).max())[1] == 0
    S = spk.Spike(partitions=4)
    S.keep_original(True)
    S.set_band_dense(a, k)
    S.set_scaling(r, c)
    with pytest.raises(spk.SpikeError):
        S.set_scaling(r, c)                                                # once per band
    S.factor()
    info = S.view()
    assert info["boosted_pivots"] == 0 and info["anorm_max"] <= 1.0 + 1e-10 if scales == "mc64" else info["boosted_pivots"] == 0
    x = S.solve(b)
    assert relerr(x, xref) < RTOL and relerr(x, u) < 1e-9
    assert relerr(S.mult(u), b) < 1e-14                                    # the kept original is the unscaled band
    S.close()


# ------------------------------------------------------------------ Krylov iteration parity (+-1)
@pytest.mark.parametrize("method", ["gmres", "bcgs"])
def test_krylov_iteration_parity(spk, oracle, method):
    n, k = 6000, 10
    A = _sparse_case(n, k, seed=21, extra=3)
    kk, f = oracle.band_select(A.indptr, A.indices, A.data, k, 1.0)
    band = oracle.csr_to_band(A.indptr, A.indices, A.data, kk)
    lu, _ = oracle.band_lu(band)
    u = np.ones(n)
    b = A @ u
    m_o = oracle.GMRES if method == "gmres" else oracle.BICGSTAB
    xo, its_o, _, rc = oracle.krylov_csr_band(A.indptr, A.indices, A.data, lu, b, m_o, rtol=1e-8)
    assert rc == 0
    S = spk.Spike(partitions=4, tip_tiles=-1)
    k2, f2 = S.set_band_csr(A.indptr, A.indices, A.data, k, 1.0)
    assert (k2, f2) == (kk, f)
    S.set_operator_csr(A.indptr, A.indices, A.data)
    S.factor()
    x, its, rn, conv = S.krylov(b, spk.GMRES if method == "gmres" else spk.BCGS, rtol=1e-8)
    assert conv and abs(its - its_o) <= 1, (its, its_o)
    assert relerr(x, u) < 1e-6 and relerr(x, xo) < 1e-6
    S.close()


@pytest.mark.parametrize("method", ["gmres", "bcgs"])
def test_krylov_runs_are_bit_identical(spk, method):
    """The inner products of the device Krylov loop are reduced in a fixed order (block partials + one finishing block,
    csrc/krylov.cu), so iteration count, residual norm and solution of repeated runs are bit-identical."""
    n, k = 20_000, 12
    A = _sparse_case(n, k, seed=33, extra=4)
    b = A @ np.linspace(0.5, 1.5, n)
    first = None
    for rep in range(4):
        S = spk.Spike(partitions=8, tip_tiles=-1)
        S.set_band_csr(A.indptr, A.indices, A.data, 6, 1.0)            # a narrower band than the matrix: a real preconditioner
        S.set_operator_csr(A.indptr, A.indices, A.data)
        S.factor()
        x, its, rn, conv = S.krylov(b, spk.GMRES if method == "gmres" else spk.BCGS, rtol=1e-10)
        S.close()
        assert conv and its > 2
        if first is None:
            first = (x.copy(), its, rn)
        else:
            assert its == first[1] and rn == first[2]
            np.testing.assert_array_equal(x, first[0])


def test_reordered_solve_end_to_end(spk, oracle):
    """testbed2 flow (src/testbed2.c:110-132) through KSPREORDER semantics (src/kspreorder.c:17-24,
    122-127): b = A u, permute operators and vectors, SPIKE-preconditioned BiCGStab, un-permute."""
    n, k = 5000, 8
    A0 = _sparse_case(n, k, seed=33, extra=1)
    rng = np.random.default_rng(7)
    q = rng.permutation(n).astype(np.int32)
    A = sp.csr_matrix(A0[q, :][:, q]); A.sort_indices()
    rorder = np.argsort(q).astype(np.int32); corder = rorder.copy()
    u = np.ones(n); b = A @ u
    S = spk.Spike(partitions=4)
    kk, f = S.set_band_csr(A.indptr, A.indices, A.data, 50, 0.95, rowperm=rorder, colperm=corder)
    assert 0 < kk <= k
    S.set_operator_csr(A.indptr, A.indices, A.data, rowperm=rorder, colperm=corder)
    S.factor()
    bp = S.permute(rorder, b, inverse=False)                     # VecPermute(b, rorder, FALSE)
    xp, its, rn, conv = S.krylov(bp, spk.BCGS, rtol=1e-10)
    x = S.permute(corder, xp, inverse=True)                      # VecPermute(x, corder, TRUE)
    assert conv and np.linalg.norm(x - u) / np.sqrt(n) < 1e-7
    S.close()


# ------------------------------------------------------------------ full-size properties (no oracle at this size)
@pytest.mark.parametrize("n,k", [(1_000_000, 50), (10_000_000, 100)])
def test_full_size_manufactured_solution(spk, n, k):
    """BASELINE configs C2 / C3 on one GPU: b = A*1 built on the device, factor, solve, ||x-1||/||1||;
    linearity: solve(2b) == 2 solve(b) bit-for-bit up to fp64 scaling exactness."""
    import torch
    S = spk.Spike(mem=spk.MEM_DEVICE)
    S.keep_original(True)
    S.set_band_synthetic(n, k)
    u = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.empty_like(u); x = torch.empty_like(u); y = torch.empty_like(u)
    S.mult(u.data_ptr(), b.data_ptr())
    S.factor()
    S.solve(b.data_ptr(), x.data_ptr())
    torch.cuda.synchronize()
    err = ((x - u).norm() / u.norm()).item()
    assert err < RTOL, err
    S.mult(x.data_ptr(), y.data_ptr())                            # residual through the kept original band
    torch.cuda.synchronize()
    assert ((y - b).norm() / b.norm()).item() < 1e-12
    b2 = 2.0 * b
    S.solve(b2.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(y, 2.0 * x)                                # scaling by 2 is exact in fp64
    assert S.view()["boosted_pivots"] == 0
    S.close()


# ------------------------------------------------------------------ the configurations bench.py times, as it times them
@pytest.mark.parametrize("delta", [1.2, 1.0])
def test_bench_config_c2_against_oracle(spk, oracle, delta):
    """BASELINE config 2 exactly as bench.py runs it (N = 1M, K = 50, 592 partitions, 48-tile truncation window):
    the whole solution against the oracle's exact no-pivot band LU solve (it fits: 0.8 GB), delta = 1.2 and 1.0."""
    import sys
    sys.path.insert(0, ".")
    from bench import CONFIGS
    cfg = CONFIGS["c2"]
    n, k = cfg["n"], cfg["k"]
    a = oracle.gen_band(n, k, 20140601, delta)
    u = oracle.gen_vec(n, 5)
    b = oracle.band_mult(a, u)
    lu, nb = oracle.band_lu(a)
    xref = oracle.band_solve(lu, b)
    S = spk.Spike(partitions=cfg["parts"], tip_tiles=cfg["tip"])
    S.set_band_synthetic(n, k, seed=20140601, delta=delta)       # the band bench.py generates (bit-identical to `a`)
    S.factor()
    x = S.solve(b)
    info = S.view()
    assert info["partitions"] == cfg["parts"] and info["tip_tiles"] == cfg["tip"] and info["boosted_pivots"] == nb == 0
    assert relerr(x, xref) < RTOL
    S.close()


@pytest.mark.parametrize("delta", [1.2, 1.0])
def test_bench_config_c3_window_and_partitions(spk, oracle, delta):
    """BASELINE config 3 exactly as bench.py runs it (N = 10M, K = 100, 296 partitions, 78-tile window): random
    manufactured solution, error and residual through the kept original; and the oracle on a 1M-row system with the
    same partition length and window (30 partitions of ~33k rows), where the exact band solve still fits."""
    import sys
    import torch
    sys.path.insert(0, ".")
    from bench import CONFIGS
    cfg = CONFIGS["c3"]
    n, k = cfg["n"], cfg["k"]
    S = spk.Spike(partitions=cfg["parts"], tip_tiles=cfg["tip"], mem=spk.MEM_DEVICE)
    S.keep_original(True)
    S.set_band_synthetic(n, k, seed=20140601, delta=delta)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.empty_like(u); x = torch.empty_like(u); y = torch.empty_like(u)
    S.mult(u.data_ptr(), b.data_ptr())
    S.factor()
    S.solve(b.data_ptr(), x.data_ptr())
    S.mult(x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    info = S.view()
    assert info["partitions"] == cfg["parts"] and info["tip_tiles"] == cfg["tip"] and info["boosted_pivots"] == 0
    assert ((x - u).norm() / u.norm()).item() < RTOL
    assert ((y - b).norm() / b.norm()).item() < 1e-11
    S.close(); del u, b, x, y
    n1 = 1_000_000
    a = oracle.gen_band(n1, k, 20140601, delta)
    u1 = oracle.gen_vec(n1, 5)
    b1 = oracle.band_mult(a, u1)
    lu, _ = oracle.band_lu(a)
    S = spk.Spike(partitions=30, tip_tiles=cfg["tip"])
    S.set_band_synthetic(n1, k, seed=20140601, delta=delta)
    S.factor()
    assert relerr(S.solve(b1), oracle.band_solve(lu, b1)) < RTOL
    S.close()


@pytest.mark.parametrize("n,k,P,nrhs", [(60_000, 100, 8, 1), (30_000, 37, 6, 9), (16_384, 256, 3, 4), (12_288, 512, 2, 1)])
def test_repeated_runs_are_bit_identical(spk, oracle, n, k, P, nrhs):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed.log), so the barrier-free dataflow of
    the LU kernels (mbarrier packages, release/acquire flags between CTAs) is checked for races the indirect way: five
    factor + solve runs of the same input must give bit-identical factors and solutions -- a missed dependency shows up
    as run-to-run differences -- and the result must agree with the oracle."""
    a = oracle.gen_band(n, k)
    U = np.stack([oracle.gen_vec(n, 20 + c) for c in range(nrhs)])
    Bm = np.stack([oracle.band_mult(a, u) for u in U])
    first = None
    for rep in range(5):
        S = spk.Spike(partitions=P, tip_tiles=0)
        S.set_band_dense(a, k)
        S.factor()
        X = S.solve(Bm if nrhs > 1 else Bm[0], nrhs=nrhs)
        F = S.get_band_rows()                                   # the factored band
        S.close()
        if first is None:
            first = (F, X)
        else:
            np.testing.assert_array_equal(F, first[0])
            np.testing.assert_array_equal(X, first[1])
    lu, _ = oracle.band_lu(a)
    X = first[1] if nrhs > 1 else first[1][None, :]
    for c in range(nrhs):
        assert relerr(X[c], oracle.band_solve(lu, Bm[c])) < RTOL


@pytest.mark.parametrize("n,k,P,nrhs", [(60_000, 100, 8, 1), (40_000, 50, 12, 1), (30_000, 37, 6, 5)])
def test_side_stream_is_transparent(spk, oracle, n, k, P, nrhs, monkeypatch):
    """The spike tips / reduced blocks of a narrow-band factorisation are queued on the context's side stream and joined
    by their first consumer (capi.cu SideScope / side_join).  Whatever the caller does between spk_factor and the
    solve -- nothing, a second factorisation from the kept original, spk_view, a second solve -- the solution is
    bit-identical to the one with SPIKE_B200_SIDE_STREAM=0 (everything on one stream) and agrees with the oracle."""
    a = oracle.gen_band(n, k)
    U = np.stack([oracle.gen_vec(n, 40 + c) for c in range(nrhs)])
    Bm = np.stack([oracle.band_mult(a, u) for u in U])
    rhs = Bm if nrhs > 1 else Bm[0]
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SPIKE_B200_SIDE_STREAM", mode)
        res = []
        for flow in ("plain", "refactor", "view", "twice"):
            S = spk.Spike(partitions=P, tip_tiles=0)
            S.keep_original(True)
            S.set_band_dense(a, k)
            S.factor()
            if flow == "refactor":
                S.factor()                      # joins the first factorisation's tips before rewriting what they read
            if flow == "view":
                assert S.view()["factored"]
            X = S.solve(rhs, nrhs=nrhs)
            if flow == "twice":
                np.testing.assert_array_equal(S.solve(rhs, nrhs=nrhs), X)
            res.append(X)
            S.close()
        for X in res[1:]:
            np.testing.assert_array_equal(X, res[0])
        out[mode] = res[0]
    np.testing.assert_array_equal(out["0"], out["1"])
    lu, _ = oracle.band_lu(a)
    X = out["1"] if nrhs > 1 else out["1"][None, :]
    for c in range(nrhs):
        assert relerr(X[c], oracle.band_solve(lu, Bm[c])) < RTOL


# ------------------------------------------------------------------ bands whose spikes do not decay
def _slow_decay_band(n):
    """shifted second difference tridiag(-1, 2.0001, -1): well conditioned (4e4), the no-pivot LU is stable, but the
    spikes of a partitioned solve decay like 0.99^rows -- truncation inside any reasonable window is not admissible."""
    a = np.zeros((n, 3))
    a[:, 0], a[:, 1], a[:, 2] = -1.0, 2.0001, -1.0
    a[0, 0] = 0; a[-1, 2] = 0
    return a


def test_self_check_flags_truncation_on_non_dominant_band(spk, oracle):
    """spk_check: on a non-dominant band several partitions give an O(1) apply error (and say so); one partition --
    the exact mode -- reproduces the reference CPU path (exact band LU)."""
    n = 40_000
    a = _slow_decay_band(n)
    lu, _ = oracle.band_lu(a)
    b = oracle.band_mult(a, oracle.gen_vec(n, 3))
    xref = oracle.band_solve(lu, b)
    S = spk.Spike(partitions=8)
    S.keep_original(True)
    S.set_band_dense(a, 1)
    S.factor()
    assert S.check() > 1e-3                      # truncated SPIKE is not an exact solve here ...
    S.close()
    S = spk.Spike(partitions=1)
    S.keep_original(True)
    S.set_band_dense(a, 1)
    S.factor()
    assert S.check() < 1e-10                     # ... one partition is
    assert relerr(S.solve(b), xref) < RTOL
    S.close()
    # and on a dominant band the check confirms the truncated solve
    S = spk.Spike(partitions=8)
    S.keep_original(True)
    S.set_band_synthetic(n, 20)
    S.factor()
    assert S.check() < 1e-12
    S.close()
