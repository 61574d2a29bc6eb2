"""GPU AWBM (spk_awbm_csr, csrc/awbm.cu) against the oracle's serial restatement of
/root/reference/src/petsc_mat_awbm.c:42-225: matching and row permutation BIT-EXACT, scalings to rounding."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden.json")


def _compare(S, oracle, A, expect_host=None):
    A = sp.csr_matrix(A)
    A.sort_indices()
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    permR, match, stats, sr, sc = S.awbm(ia, ja, a, want_scalings=True)
    o_perm, _, o_match = oracle.awbm(ia, ja, a)
    np.testing.assert_array_equal(match, o_match)
    np.testing.assert_array_equal(permR, o_perm)
    assert sorted(permR.tolist()) == list(range(A.shape[0]))
    if expect_host is not None:
        assert (stats[2] + stats[3] > 0) == expect_host, stats
    return stats, sr, sc


def test_awbm_known_answer_of_the_reference(spk, oracle):
    """the 3x3 matrix of src/wbm.c:483-497 (tests/golden/golden.json)"""
    g = json.load(open(GOLDEN))["wbm3x3"]
    S = spk.Spike()
    permR, match, stats = S.awbm(np.array(g["ia"]), np.array(g["ja"]), np.array(g["a"]))
    assert match.tolist() == g["awbm_match"] == [1, 2, 0]
    assert permR.tolist() == g["awbm_permR"] == [2, 0, 1]


@pytest.mark.parametrize("n,density,seed", [(50, 0.2, 1), (500, 0.02, 2), (5000, 0.002, 3), (60000, 0.0002, 4)])
def test_awbm_random_sparse(spk, oracle, n, density, seed):
    """unstructured random matrices: many columns compete for the same tight rows, some need the repair passes"""
    rng = np.random.default_rng(seed)
    m = int(density * n * n)
    A = sp.csr_matrix((rng.uniform(-1, 1, m), (rng.integers(0, n, m), rng.integers(0, n, m))), shape=(n, n))
    A = A + sp.diags(rng.uniform(-1, 1, n) * (rng.random(n) < 0.7))      # 30 % of the diagonal is structurally absent
    A = sp.csr_matrix(A); A.sum_duplicates()
    S = spk.Spike()
    _compare(S, oracle, A)


def test_awbm_scrambled_dominant_matrix_and_scalings(spk, oracle):
    """the testbed2 use: dominant entries moved off the diagonal by a row permutation the matching must undo
    (tests/test_gpu_petsc_glue.py); scalings exp(v)/amax, exp(u) of src/petsc_mat_awbm.c:212-215 make every matched
    entry 1 in magnitude and no entry larger"""
    n, k = 200000, 6
    rng = np.random.default_rng(7)
    diags = [rng.uniform(-1, 1, n) for _ in range(2 * k + 1)]
    diags[k] = 3.0 * k + rng.uniform(0, 1, n)
    A0 = sp.diags(diags, list(range(-k, k + 1)), shape=(n, n), format="csr")
    R = rng.permutation(n)
    A = sp.csr_matrix(A0[R, :])
    S = spk.Spike()
    stats, sr, sc = _compare(S, oracle, A)
    assert stats[1] == n and stats[0] <= 4          # everything matched on the device in a few rounds
    A.sort_indices()
    permR, match, _ = S.awbm(A.indptr, A.indices, A.data)
    # CSR row c is matched to its dominant entry (column of A0's diagonal = R[c])
    np.testing.assert_array_equal(match, R)
    Sc = sp.diags(sr) @ abs(A) @ sp.diags(sc)
    assert Sc.max() <= 1.0 + 1e-12
    np.testing.assert_allclose(np.asarray(Sc[np.arange(n), match]).ravel(), 1.0, rtol=1e-12)


def test_awbm_chain_hands_over_to_the_serial_rule(spk, oracle):
    """bidiagonal with equal magnitudes, listed sub-diagonal first: every column's first tight row is wanted by its
    predecessor -- one link per round on the device, so the remainder is finished by the serial rule on the host"""
    n = 20000
    A = sp.diags([np.ones(n - 1), np.ones(n)], [-1, 0], format="csr")
    S = spk.Spike()
    stats, _, _ = _compare(S, oracle, A, expect_host=True)
    assert stats[0] < 64


def test_awbm_zero_entries_and_empty_rows(spk, oracle):
    """explicit zeros get weight DBL_MAX (:77), rows without entries fall to the completion pass (:181-193)"""
    rng = np.random.default_rng(11)
    n = 3000
    m = int(0.003 * n * n)
    A = sp.csr_matrix((rng.uniform(0.1, 1, m), (rng.integers(0, n, m), rng.integers(0, n, m))), shape=(n, n)).tolil()
    for r in rng.choice(n, 40, replace=False):
        A.rows[r], A.data[r] = [], []
    A = sp.csr_matrix(A)
    A.data[rng.choice(A.nnz, A.nnz // 10, replace=False)] = 0.0          # explicit zeros stay in the pattern
    S = spk.Spike()
    _compare(S, oracle, A, expect_host=True)


def test_awbm_ordering_registered_in_the_glue(spk, oracle):
    """MatGetOrdering_AWBM of the host glue (host/ordering.c) returns the reference's ISs: row = permR, col = identity"""
    L = C.CDLL(os.path.join(os.path.dirname(spk.library_path()), "libspike_petsc.so"))
    n = 4000
    rng = np.random.default_rng(3)
    A0 = sp.diags([rng.uniform(-1, 1, n), 5 + rng.random(n), rng.uniform(-1, 1, n)], [-2, 0, 3], shape=(n, n), format="csr")
    A = sp.csr_matrix(A0[rng.permutation(n), :]); A.sort_indices()
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    class ISS(C.Structure):
        _fields_ = [("n", C.c_int), ("idx", C.POINTER(C.c_int))]
    m = C.c_void_p()
    L.MatCreateSeqAIJWithArrays.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    assert L.MatCreateSeqAIJWithArrays(n, ia.ctypes.data, ja.ctypes.data, a.ctypes.data, C.byref(m)) == 0
    row, col = C.c_void_p(), C.c_void_p()
    L.MatGetOrdering_AWBM.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.PetscLastErrorMessage.restype = C.c_char_p
    assert L.MatGetOrdering_AWBM(m, b"awbm", C.byref(row), C.byref(col)) == 0, L.PetscLastErrorMessage()
    rs, cs = C.cast(row, C.POINTER(ISS)).contents, C.cast(col, C.POINTER(ISS)).contents
    assert rs.n == n and cs.n == n
    pr = np.ctypeslib.as_array(rs.idx, (n,)).copy()
    pc = np.ctypeslib.as_array(cs.idx, (n,)).copy()
    o_perm, o_col, _ = oracle.awbm(ia, ja, a)
    np.testing.assert_array_equal(pr, o_perm)
    np.testing.assert_array_equal(pc, o_col)
