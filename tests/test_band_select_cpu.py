"""Host part of spk_set_band_csr (csrc/capi.cu: band_select_host -- multi-threaded gather + one ordered summation pass)
against the oracle's restatement of MatPermute + MatCreateSubMatrixBanded (/root/reference/src/kspreorder.c:20,
src/matbanded.c:38-56,104-105): k and the norm fraction BIT-EXACT, including the fall-through quirk and the
out-of-bounds case.  Host-only code path: runs without a GPU."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp


def _ours(L, ia, ja, a, rp, cp, kmax, frac):
    k, f = C.c_int(), C.c_double()
    rc = L.spk_debug_band_select(len(ia) - 1, ia.ctypes.data, ja.ctypes.data, a.ctypes.data,
                                 rp.ctypes.data if rp is not None else None, cp.ctypes.data if cp is not None else None,
                                 kmax, frac, C.byref(k), C.byref(f))
    return rc, k.value, f.value


@pytest.mark.parametrize("n,per_row,seed", [(7, 3, 0), (300, 6, 1), (5000, 12, 2), (200000, 9, 3), (2000, 150, 4)])
def test_band_select_host_is_bit_exact(spk, oracle, n, per_row, seed):
    L = spk.lib()
    L.spk_debug_band_select.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), per_row)
    cols = rng.integers(0, n, rows.size)
    A = sp.csr_matrix((rng.uniform(-1, 1, rows.size), (rows, cols)), shape=(n, n)) + sp.eye(n) * rng.uniform(0.1, 3)
    A = sp.csr_matrix(A); A.sum_duplicates(); A.sort_indices()
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    rp = rng.permutation(n).astype(np.int32)
    cp = rng.permutation(n).astype(np.int32)
    ident = np.arange(n, dtype=np.int32)
    for r, c in [(None, None), (rp, cp), (rp, None), (None, cp), (rp, rp)]:
        ib, jb, b = oracle.mat_permute_csr(ia, ja, a, r if r is not None else ident, c if c is not None else ident)
        for kmax, frac in [(5, 0.5), (50, 0.95), (n + 3, 1.0), (1, 0.1), (n, 2.0)]:
            try:
                ko, fo = oracle.band_select(ib, jb, b, kmax, frac)
                bad = False
            except ValueError:
                bad = True
            rc, k, f = _ours(L, ia, ja, a, r, c, kmax, frac)
            assert (rc != 0) == bad
            if not bad:
                assert k == ko and f == fo          # bit-exact (== on doubles)
