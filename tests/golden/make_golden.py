"""Regenerates tests/golden/*.json from the reference's own inputs.

Run in the build container (needs /root/reference for MC64):  python tests/golden/make_golden.py
Sources of the vectors:
  * 3x3 matrix "From HC64 documentation"  -- /root/reference/src/wbm.c:483-497 (input only; the
    reference records no expected output).  Outputs below come from (a) the reference's own MC64
    (src/hslmc64.c compiled as-is, called with the convention of src/petsc_mat_wbm.c:20-58) and
    (b) the AWBM restatement (src/petsc_mat_awbm.c:65-205); both equal the probe answers recorded in
    SURVEY.md section 8c.
  * band selector table on T = pentadiag(.5,1,4,1,.5) 4x4 -- SURVEY.md section 8c probe of
    src/matbanded.c:38-56,104-105.
  * synthetic generator / exact banded solve vectors -- produced by the oracle and cross-checked
    against scipy.linalg.solve_banded (LAPACK dgbsv) at generation time.
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sl
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def main():
    out = {}
    A = np.array([[0, 8, 3], [0, 2, 1], [4, 0, 0]], float)
    M = sp.csr_matrix(A)
    row, col, num, dw = O.wbm(M.indptr, M.indices, M.data)
    pr, pc, match = O.awbm(M.indptr, M.indices, M.data)
    out["wbm3x3"] = {"ia": M.indptr.tolist(), "ja": M.indices.tolist(), "a": M.data.tolist(),
                     "mc64_perm_1based": (col + 1).tolist(), "num": int(num), "dw": dw[:6].tolist(),
                     "row_is": row.tolist(), "col_is": col.tolist(),
                     "awbm_match": match.tolist(), "awbm_permR": pr.tolist(), "awbm_permC": pc.tolist(),
                     "awbm_diag": A[pr, :].diagonal().tolist()}
    T = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            T[i, j] = {0: 4.0, 1: 1.0, 2: 0.5}.get(abs(i - j), 0.0)
    Ts = sp.csr_matrix(T)
    sel = []
    for kmax, frac in [(50, .95), (1, .95), (50, .8), (2, 1.0), (3, .5)]:
        k, f = O.band_select(Ts.indptr, Ts.indices, Ts.data, kmax, frac)
        sel.append({"kmax": kmax, "frac": frac, "k": k, "frac_out": f})
    out["band_select_penta4"] = {"ia": Ts.indptr.tolist(), "ja": Ts.indices.tolist(), "a": Ts.data.tolist(), "cases": sel}
    # generator spot values + exact solve fixture
    n, k = 64, 5
    a = O.gen_band(n, k)
    u = O.gen_vec(n)
    b = O.band_mult(a, u)
    lu, nb = O.band_lu(a)
    x = O.band_solve(lu, b)
    ab = np.zeros((2 * k + 1, n))
    for d in range(-k, k + 1):
        i0, i1 = max(0, -d), min(n, n - d)
        ab[k - d, i0 + d:i1 + d] = a[i0:i1, d + k]
    xs = sl.solve_banded((k, k), ab, b)
    assert np.abs(xs - x).max() < 1e-13
    out["synthetic_n64_k5"] = {"seed": 20140601, "delta": 1.2, "band_row0": a[0].tolist(), "band_row17": a[17].tolist(),
                               "band_row63": a[63].tolist(), "u_first8": u[:8].tolist(), "b_first8": b[:8].tolist(),
                               "x_first8": x[:8].tolist(), "lu_row17": lu[17].tolist(),
                               "u01_samples": [O.u01(20140601, c) for c in (0, 1, 2, 12345678901)]}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
