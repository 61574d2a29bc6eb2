"""BASELINE config 4 end to end: synthetic sparse nonsymmetric matrix (band pattern scrambled by a symmetric
permutation Q and a row permutation R) -> WBM (the reference's own MC64, host callback) -> "fiedler" (MC73 is absent:
the deterministic RCM stand-in of host/ordering.c) -> PCBANDED(kmax, 0.95) -> SPIKE-preconditioned BiCGStab, rtol 1e-5
(src/makefile:18), driven through the C glue exactly like src/testbed2.c:61-73,110-132 drives the reference:
types registered by name, everything selected by prefixed options, manufactured solution u = 1.

Bars (north_star): k / frac / band structure bit-exact vs the oracle's restatement of MatPermute +
MatCreateSubMatrixBanded; Krylov iterations within +-1 of the oracle's BiCGStab with the exact band solve; ||x - u||.
"""
import ctypes as C
import os
import time

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "spike_petsc_b200", "lib", "libspike_petsc.so")
ORDFN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))


class MatS(C.Structure):
    _fields_ = [("n", C.c_int), ("i", C.POINTER(C.c_int)), ("j", C.POINTER(C.c_int)), ("a", C.POINTER(C.c_double)), ("refct", C.c_int)]


def _glue():
    L = C.CDLL(LIB)
    L.PetscLastErrorMessage.restype = C.c_char_p
    L.MatCreateSeqAIJWithArrays.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.VecCreateSeqWithArray.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.ISCreateGeneral.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.MatOrderingRegister.argtypes = [C.c_char_p, ORDFN]
    L.PetscOptionsSetValue.argtypes = [C.c_char_p, C.c_char_p]
    L.KSPCreate.argtypes = [C.POINTER(C.c_void_p)]
    for f in ("KSPCreate_Reorder", "KSPSetFromOptions"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.KSPSetOperators.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.KSPSolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.KSPGetIterationNumber.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.KSPGetConvergedReason.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.KSPDestroy.argtypes = [C.POINTER(C.c_void_p)]
    L.KSPView.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.SpkOrderingRCM.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def two_stage_ordering(L, oracle, ia, ja, a):
    """-mat_ordering_type wbm, then -mat_ordering_type2 fiedler on the permuted matrix (src/testbed.c:200-284), folded
    into one (row, col) pair for KSPREORDER."""
    from spike_petsc_b200 import synthetic
    n = len(ia) - 1
    _, c1, num, _ = oracle.wbm(ia, ja, a)
    assert num == n
    PM = sp.csr_matrix(sp.csr_matrix((a, ja, ia), shape=(n, n))[:, c1])
    PM.sort_indices()
    pi, pj = PM.indptr.astype(np.int32), PM.indices.astype(np.int32)
    p2 = np.zeros(n, dtype=np.int32)
    assert L.SpkOrderingRCM(n, pi.ctypes.data, pj.ctypes.data, p2.ctypes.data) == 0
    return synthetic.compose_wbm_then_symmetric(c1, p2)


@pytest.mark.parametrize("n,hb", [(100_000, 300), (2_000_000, 2000)])
def test_c4_wbm_fiedler_pcbanded_bicgstab(spk, oracle, n, hb):
    from spike_petsc_b200 import synthetic
    if not oracle.have_mc64():
        pytest.skip("oracle/_ref (the reference's MC64) not built")
    L = _glue()
    A, Q, R = synthetic.c4_matrix(n, hb)
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    timing = {}
    perms = {}

    def cb(mat, typ, row, col):
        Ms = C.cast(mat, C.POINTER(MatS)).contents
        mi = np.ctypeslib.as_array(Ms.i, (Ms.n + 1,)); mj = np.ctypeslib.as_array(Ms.j, (mi[-1],)); ma = np.ctypeslib.as_array(Ms.a, (mi[-1],))
        t0 = time.perf_counter()
        rp, cp = two_stage_ordering(L, oracle, mi, mj, ma)
        timing["ordering_s"] = time.perf_counter() - t0
        perms["row"], perms["col"] = rp, cp
        L.ISCreateGeneral(Ms.n, rp.ctypes.data, C.cast(row, C.POINTER(C.c_void_p)))
        L.ISCreateGeneral(Ms.n, cp.ctypes.data, C.cast(col, C.POINTER(C.c_void_p)))
        return 0

    fn = ORDFN(cb)
    L.MatOrderingRegister(b"wbm+fiedler", fn)
    L.PetscOptionsClear()
    for name, val in [("-mat_ordering_type", "wbm+fiedler"), ("-reorder_ksp_type", "bcgs"), ("-reorder_pc_type", "banded"),
                      ("-reorder_ksp_rtol", "1e-5"), ("-reorder_ksp_max_it", "500"),
                      ("-reorder_pc_banded_kmax", "50"), ("-reorder_pc_banded_frac", "0.95")]:
        L.PetscOptionsSetValue(name.encode(), val.encode())
    m = C.c_void_p()
    assert L.MatCreateSeqAIJWithArrays(n, ia.ctypes.data, ja.ctypes.data, a.ctypes.data, C.byref(m)) == 0
    ksp = C.c_void_p(); L.KSPCreate(C.byref(ksp)); L.KSPCreate_Reorder(ksp)
    L.KSPSetOperators(ksp, m, m)
    assert L.KSPSetFromOptions(ksp) == 0, L.PetscLastErrorMessage()
    u = np.ones(n); b = np.ascontiguousarray(A @ u); b0 = b.copy(); x = np.zeros(n)
    vb, vx = C.c_void_p(), C.c_void_p()
    L.VecCreateSeqWithArray(n, b.ctypes.data, C.byref(vb)); L.VecCreateSeqWithArray(n, x.ctypes.data, C.byref(vx))
    t0 = time.perf_counter()
    assert L.KSPSolve(ksp, vb, vx) == 0, L.PetscLastErrorMessage()
    timing["kspsolve_s"] = time.perf_counter() - t0
    reason, its = C.c_int(), C.c_int()
    L.KSPGetConvergedReason(ksp, C.byref(reason)); L.KSPGetIterationNumber(ksp, C.byref(its))
    buf = C.create_string_buffer(1024); L.KSPView(ksp, buf, 1024)
    view = buf.value.decode()
    np.testing.assert_array_equal(b, b0)             # permuted in place and restored (src/kspreorder.c:123,127)
    # ---- the oracle on the same inputs: MatPermute, band selection, exact band LU, BiCGStab
    rp, cp = perms["row"], perms["col"]
    pia, pja, pa = oracle.mat_permute_csr(ia, ja, a, rp, cp)
    kref, fref = oracle.band_select(pia, pja, pa, 50, 0.95)
    assert view.startswith("  reordering type = wbm+fiedler\n  Banded: k = %d (50 max), frac = " % kref), view
    assert ("frac = %g (0.95 max)" % fref) in view     # PCView_Banded prints %g (src/matbanded.c:205)
    band_ref = oracle.csr_to_band(pia, pja, pa, kref)
    lu, nboost = oracle.band_lu(band_ref)
    bp = b0[rp]
    xo, its_ref, _, rc = oracle.krylov_csr_band(pia, pja, pa, lu, bp, method=oracle.BICGSTAB, rtol=1e-5, maxit=500)
    assert rc == 0 and reason.value > 0
    assert abs(its.value - its_ref) <= 1, (its.value, its_ref)
    err = np.linalg.norm(x - u) / np.sqrt(n)          # "Error in solution" of src/testbed2.c:130-132, per entry
    assert err < 1e-4, err
    # ---- band structure / values bit-exact: the same fused permutation + extraction on its own context
    S = spk.Spike()
    t0 = time.perf_counter()
    k, f = S.set_band_csr(ia, ja, a, 50, 0.95, rowperm=rp, colperm=cp)
    timing["band_select_and_pack_s"] = time.perf_counter() - t0
    assert (k, f) == (kref, fref)
    np.testing.assert_array_equal(S.get_band_rows(), band_ref)
    S.close()
    print(f"\nC4 n={n}: its {its.value} (oracle {its_ref}), err/entry {err:.2e}, k={kref} frac={fref:.4f}, "
          f"ordering {timing['ordering_s']:.1f}s, KSPSolve (setup+solve) {timing['kspsolve_s']:.2f}s, "
          f"select+pack {timing['band_select_and_pack_s']:.2f}s")
    L.KSPDestroy(C.byref(ksp))
