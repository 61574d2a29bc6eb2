"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors / KATs
 and against independent LAPACK results, host-side logic, and that the C-ABI
library loads and exports every symbol include/spike_b200.h declares."""
import ctypes
import json
import os

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def _scipy_band(a, k):
    n = a.shape[0]
    ab = np.zeros((2 * k + 1, n))
    for d in range(-k, k + 1):
        i0, i1 = max(0, -d), min(n, n - d)
        ab[k - d, i0 + d:i1 + d] = a[i0:i1, d + k]
    return ab


# ---------------------------------------------------------------- golden vectors
def test_mc64_reference_kat(oracle):
    """Reference's own MC64 (src/hslmc64.c built as-is) on the 3x3 matrix of src/wbm.c:483-497,
    through the wrapper convention of src/petsc_mat_wbm.c:20-58."""
    if not oracle.have_mc64():
        pytest.skip("oracle/_ref/libmc64ref.so not built (reference tree absent)")
    g = GOLD["wbm3x3"]
    row, col, num, dw = oracle.wbm(g["ia"], g["ja"], g["a"])
    assert num == 3
    assert (col + 1).tolist() == [3, 1, 2] == g["mc64_perm_1based"]
    assert row.tolist() == [0, 1, 2]
    np.testing.assert_allclose(dw[:6], [0, 0, np.log(2), -np.log(8), -np.log(2), -np.log(4)], rtol=0, atol=1e-15)
    np.testing.assert_array_equal(dw[:6], g["dw"])
    # matched entries are A(perm[i], i): product 4*8*1 = 32
    A = sp.csr_matrix((g["a"], g["ja"], g["ia"]), shape=(3, 3)).toarray()
    assert np.prod([A[col[i], i] for i in range(3)]) == 32.0


def test_awbm_kat(oracle):
    g = GOLD["wbm3x3"]
    permR, permC, match = oracle.awbm(g["ia"], g["ja"], g["a"])
    assert match.tolist() == [1, 2, 0] == g["awbm_match"]
    assert permR.tolist() == [2, 0, 1] == g["awbm_permR"]
    assert permC.tolist() == [0, 1, 2]
    A = sp.csr_matrix((g["a"], g["ja"], g["ia"]), shape=(3, 3)).toarray()
    assert A[permR, :].diagonal().tolist() == [4.0, 8.0, 1.0]


def test_band_select_kat(oracle):
    g = GOLD["band_select_penta4"]
    expect = {(50, .95): (2, 1.0), (1, .95): (1, 2 / 3), (50, .8): (1, 11 / 12), (2, 1.0): (2, 11 / 12), (3, .5): (0, 2 / 3)}
    for case in g["cases"]:
        k, f = oracle.band_select(g["ia"], g["ja"], g["a"], case["kmax"], case["frac"])
        assert (k, f) == (case["k"], case["frac_out"])
        ek, ef = expect[(case["kmax"], case["frac"])]
        assert k == ek and abs(f - ef) < 1e-15


def test_band_select_out_of_bounds_is_refused(oracle):
    g = GOLD["band_select_penta4"]
    with pytest.raises(ValueError):
        # frac > 1 never breaks: the reference walks past the weight vector (src/matbanded.c:54)
        oracle.band_select(g["ia"], g["ja"], g["a"], 50, 1.5)


def test_generator_golden(oracle):
    g = GOLD["synthetic_n64_k5"]
    a = oracle.gen_band(64, 5, g["seed"], g["delta"])
    np.testing.assert_array_equal(a[0], g["band_row0"])
    np.testing.assert_array_equal(a[17], g["band_row17"])
    np.testing.assert_array_equal(a[63], g["band_row63"])
    assert [oracle.u01(g["seed"], c) for c in (0, 1, 2, 12345678901)] == g["u01_samples"]
    u = oracle.gen_vec(64, g["seed"])
    np.testing.assert_array_equal(u[:8], g["u_first8"])
    lu, nb = oracle.band_lu(a)
    assert nb == 0
    np.testing.assert_array_equal(lu[17], g["lu_row17"])
    x = oracle.band_solve(lu, oracle.band_mult(a, u))
    np.testing.assert_array_equal(x[:8], g["x_first8"])
    # diagonal dominance with delta = 1.2 and out-of-range entries exactly zero
    assert np.all(a[:, 5] >= 1.2 * (np.abs(a).sum(1) - a[:, 5]) * (1 - 1e-15))
    assert np.all(a[0, :5] == 0) and np.all(a[63, 6:] == 0)


# ---------------------------------------------------------------- exact banded solve vs LAPACK
@pytest.mark.parametrize("n,k", [(1, 0), (7, 2), (200, 10), (1000, 37), (513, 100)])
def test_band_lu_vs_lapack(oracle, n, k):
    a = oracle.gen_band(n, k, seed=7 + n)
    u = oracle.gen_vec(n, 3)
    b = oracle.band_mult(a, u)
    lu, nb = oracle.band_lu(a)
    x = oracle.band_solve(lu, b)
    xs = sl.solve_banded((k, k), _scipy_band(a, k), b) if n > 1 else b / a[:, k]
    assert np.abs(x - xs).max() <= 1e-12 * max(1.0, np.abs(xs).max())
    assert np.abs(x - u).max() <= 1e-12


def test_band_lu_boosting(oracle):
    n, k = 50, 3
    a = oracle.gen_band(n, k)
    a[10, k] = 0.0
    a[10, :k] = 0.0  # force an exactly-zero pivot at row 10 (nothing above contributes)
    a[:, :] = np.where(np.abs(a) < 1e-300, 0.0, a)
    a[7:10, k + 1:] *= 1.0
    for i in range(7, 10):
        a[i, (10 - i) + k] = 0.0  # column 10 entries above the diagonal
    lu, nb = oracle.band_lu(a, boost=1e-8)
    assert nb >= 1 and np.all(np.isfinite(lu))


# ---------------------------------------------------------------- SPIKE algebra
@pytest.mark.parametrize("n,k,P", [(2000, 10, 4), (3000, 25, 7), (4096, 50, 5)])
def test_cpu_spike_matches_exact(oracle, n, k, P):
    a = oracle.gen_band(n, k)
    u = np.ones(n)
    b = oracle.band_mult(a, u)
    lu, _ = oracle.band_lu(a)
    xe = oracle.band_solve(lu, b)
    S = oracle.Spike(n, k, P, align=8, tip_rows=0)
    S.factor(a)
    x = S.solve(b)
    assert np.abs(x - xe).max() <= 1e-12
    # windowed tips / truncated corrections converge to the exact answer as the window grows
    errs = []
    for tip in (4 * k, 8 * k, 16 * k):
        St = oracle.Spike(n, k, P, align=8, tip_rows=tip)
        St.factor(a)
        errs.append(np.abs(St.solve(b) - xe).max())
    assert errs[-1] <= 1e-12 and errs[0] >= errs[-1]


def test_spike_tips_definition(oracle):
    """V^(b), W^(t) equal the tips of A_i^{-1}[0;B_i], A_i^{-1}[C_i;0] computed densely."""
    n, k, P = 600, 6, 3
    a = oracle.gen_band(n, k, seed=11)
    S = oracle.Spike(n, k, P, align=8)
    S.factor(a)
    A = np.zeros((n, n))
    for i in range(n):
        for d in range(-k, k + 1):
            if 0 <= i + d < n:
                A[i, i + d] = a[i, d + k]
    for i in range(P - 1):
        lo, mid, hi = S.part_start(i), S.part_start(i + 1), S.part_start(i + 2)
        Ai = A[lo:mid, lo:mid]
        rhs = np.zeros((mid - lo, k)); rhs[-k:, :] = A[mid - k:mid, mid:mid + k]
        np.testing.assert_allclose(S.vb(i), np.linalg.solve(Ai, rhs)[-k:], rtol=0, atol=1e-13)
        Aj = A[mid:hi, mid:hi]
        rhs = np.zeros((hi - mid, k)); rhs[:k, :] = A[mid:mid + k, mid - k:mid]
        np.testing.assert_allclose(S.wt(i), np.linalg.solve(Aj, rhs)[:k], rtol=0, atol=1e-13)


# ---------------------------------------------------------------- PETSc semantics restated
def test_mat_permute_and_vec_permute(oracle):
    rng = np.random.default_rng(5)
    n = 40
    A = sp.random(n, n, 0.2, random_state=3, format="csr") + sp.eye(n, format="csr")
    A = sp.csr_matrix(A); A.sort_indices()
    rp, cp = rng.permutation(n).astype(np.int32), rng.permutation(n).astype(np.int32)
    ib, jb, b = oracle.mat_permute_csr(A.indptr, A.indices, A.data, rp, cp)
    B = sp.csr_matrix((b, jb, ib), shape=(n, n)).toarray()
    np.testing.assert_array_equal(B, A.toarray()[np.ix_(rp, cp)])     # B(i,j) = A(rowp[i], colp[j])
    assert all(np.all(np.diff(jb[ib[i]:ib[i + 1]]) > 0) for i in range(n))
    x = rng.standard_normal(n)
    y = oracle.vec_permute(x, rp, False)
    np.testing.assert_array_equal(y, x[rp])
    np.testing.assert_array_equal(oracle.vec_permute(y, rp, True), x)
    # KSPSolve_Reorder identity (src/kspreorder.c:19-24,122-127): (P_r A P_c)(P_c^T x) = P_r b
    np.testing.assert_allclose(B @ oracle.vec_permute(x, cp, False), oracle.vec_permute(A @ x, rp, False), atol=1e-12)


def test_band_extract_preserves_order_and_values(oracle):
    A = sp.random(60, 60, 0.3, random_state=9, format="csr") + sp.eye(60, format="csr")
    A = sp.csr_matrix(A); A.sort_indices()
    ib, jb, b = oracle.band_extract_csr(A.indptr, A.indices, A.data, 4)
    D = A.toarray()
    mask = np.abs(np.subtract.outer(np.arange(60), np.arange(60))) <= 4
    np.testing.assert_array_equal(sp.csr_matrix((b, jb, ib), shape=(60, 60)).toarray(), D * mask)
    np.testing.assert_array_equal(oracle.csr_to_band(A.indptr, A.indices, A.data, 4),
                                  np.array([[D[i, i + d] if 0 <= i + d < 60 else 0.0 for d in range(-4, 5)] for i in range(60)]))


def test_awbm_random_is_a_permutation_with_nonzero_diagonal(oracle):
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = 30
        A = sp.random(n, n, 0.15, random_state=trial, format="csr") + sp.csr_matrix((np.ones(n), (rng.permutation(n), np.arange(n))), shape=(n, n))
        A = sp.csr_matrix(A); A.sort_indices()
        permR, permC, match = oracle.awbm(A.indptr, A.indices, A.data)
        assert sorted(permR.tolist()) == list(range(n))
        # the reference builds permR[match[c]] = c (src/petsc_mat_awbm.c:201)
        assert all(permR[match[c]] == c for c in range(n))


def test_mc64_random_optimal(oracle):
    if not oracle.have_mc64():
        pytest.skip("oracle/_ref not built")
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(1)
    for trial in range(10):
        n = 12
        D = rng.uniform(0.1, 1.0, (n, n)) * (rng.uniform(size=(n, n)) < 0.5)
        D[rng.permutation(n), np.arange(n)] = rng.uniform(0.5, 1.0, n)
        D = np.maximum(D, D.T * (D == 0))  # structurally symmetric pattern (MatGetRowIJ symmetric=TRUE hazard)
        A = sp.csr_matrix(D); A.sort_indices()
        row, col, num, dw = oracle.wbm(A.indptr, A.indices, A.data)
        assert num == n and sorted(col.tolist()) == list(range(n))
        got = np.sum(np.log(np.abs([D[col[i], i] for i in range(n)])))
        cost = np.where(D != 0, -np.log(np.abs(np.where(D != 0, D, 1.0))), 1e6)
        r, c = linear_sum_assignment(cost)
        assert abs(got + cost[r, c].sum()) < 1e-9


# ---------------------------------------------------------------- Krylov restatement
def test_krylov_exact_preconditioner_converges_in_one(oracle):
    a = oracle.gen_band(3000, 10)
    lu, _ = oracle.band_lu(a)
    b = oracle.band_mult(a, np.ones(3000))
    for m in (oracle.GMRES, oracle.BICGSTAB):
        x, its, rn, rc = oracle.krylov_band(a, lu, b, m)
        assert rc == 0 and its == 1 and np.abs(x - 1).max() < 1e-9


def test_krylov_band_preconditioner_on_sparse(oracle):
    n, k = 1500, 8
    a = oracle.gen_band(n, k, delta=1.05)
    rng = np.random.default_rng(2)
    D = sp.lil_matrix((n, n))
    for i in range(n):
        for d in range(-k, k + 1):
            if 0 <= i + d < n:
                D[i, i + d] = a[i, d + k]
    for _ in range(3 * n):  # weak off-band entries the band preconditioner ignores
        i, j = rng.integers(0, n, 2)
        if abs(i - j) > k:
            D[i, j] = 0.05 * rng.standard_normal()
    A = sp.csr_matrix(D); A.sort_indices()
    kk, f = oracle.band_select(A.indptr, A.indices, A.data, k, 1.0)
    assert kk == k and f < 1.0  # fall-through quirk: k == kmax, fraction excludes w[kmax]
    band = oracle.csr_to_band(A.indptr, A.indices, A.data, kk)
    lu, _ = oracle.band_lu(band)
    u = np.ones(n); b = A @ u
    xg, itg, _, rcg = oracle.krylov_csr_band(A.indptr, A.indices, A.data, lu, b, oracle.GMRES, rtol=1e-8)
    xb, itb, _, rcb = oracle.krylov_csr_band(A.indptr, A.indices, A.data, lu, b, oracle.BICGSTAB, rtol=1e-8)
    assert rcg == 0 and rcb == 0 and 1 < itg < 60 and 1 <= itb < 60
    assert np.abs(xg - u).max() < 1e-5 and np.abs(xb - u).max() < 1e-5


# ---------------------------------------------------------------- C ABI: loads + exports (no compute)
def test_cabi_exports_every_declared_symbol():
    import spike_petsc_b200 as spk
    path = spk.library_path()
    assert os.path.exists(path), "libspike_b200.so missing: run `make` / __graft_entry__.build()"
    L = ctypes.CDLL(path)
    names = spk.exported_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/spike_b200.h but not exported"
    assert b"sm_100a" in spk.lib().spk_version()


def test_cabi_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import spike_petsc_b200 as spk
    with pytest.raises(spk.SpikeError, match="no CPU fallback|CUDA"):
        spk.Spike()


def test_petsc_glue_library_exports_plugin_surface():
    """libspike_petsc.so (C host glue) loads and exports the reference's plugin entry points."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spike_petsc_b200", "lib", "libspike_petsc.so")
    assert os.path.exists(path), "libspike_petsc.so missing: run `make`"
    L = ctypes.CDLL(path)
    for name in ("PCCreate_Banded", "KSPCreate_Reorder", "MatCreateSubMatrixBanded", "PCBandedSetMaxHalfBandwith",
                 "PCBandedSetNormFraction", "MatOrderingRegister", "KSPSolve", "PCApply"):
        assert hasattr(L, name)
