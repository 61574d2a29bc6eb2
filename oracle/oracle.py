"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (spike_petsc_b200) never does.

Functions mirror oracle/spike_oracle.h; the WBM ordering calls the reference's own MC64
(/root/reference/src/hslmc64.c compiled as-is into oracle/_ref/libmc64ref.so) with the calling
convention of /root/reference/src/petsc_mat_wbm.c:20-58.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_MC64 = None

c_i64 = C.c_int64
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spike_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    need_ref = os.path.exists("/root/reference/src/hslmc64.c") and not os.path.exists(
        os.path.join(_HERE, "_ref", "libmc64ref.so"))
    if force or stale or need_ref:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_gen_band.argtypes = [c_i64, C.c_int, C.c_uint64, C.c_double, c_dp]
        L.orc_gen_vec.argtypes = [c_i64, C.c_uint64, c_dp]
        L.orc_band_select.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.c_int, C.c_double, c_ip, c_dp]
        L.orc_band_extract_csr.restype = c_i64
        L.orc_band_extract_csr.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.c_int, c_ip, c_ip, c_dp]
        L.orc_csr_to_band.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.c_int, c_dp]
        L.orc_mat_permute_csr.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_ip, c_ip, c_ip, c_ip, c_dp]
        L.orc_vec_permute.argtypes = [C.c_int, c_dp, c_ip, C.c_int]
        L.orc_awbm.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_ip, c_ip]
        L.orc_band_lu.restype = c_i64
        L.orc_band_lu.argtypes = [c_i64, C.c_int, c_dp, C.c_double]
        L.orc_band_solve.argtypes = [c_i64, C.c_int, c_dp, c_dp, C.c_int, c_i64]
        L.orc_band_mult.argtypes = [c_i64, C.c_int, c_dp, c_dp, c_dp]
        L.orc_spike_create.restype = C.c_void_p
        L.orc_spike_create.argtypes = [c_i64, C.c_int, C.c_int, C.c_int, c_i64, C.c_double]
        L.orc_spike_destroy.argtypes = [C.c_void_p]
        L.orc_spike_factor.restype = c_i64
        L.orc_spike_factor.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.orc_spike_solve.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int]
        L.orc_spike_vb.restype = c_dp
        L.orc_spike_vb.argtypes = [C.c_void_p, C.c_int]
        L.orc_spike_wt.restype = c_dp
        L.orc_spike_wt.argtypes = [C.c_void_p, C.c_int]
        L.orc_spike_part_start.restype = c_i64
        L.orc_spike_part_start.argtypes = [C.c_void_p, C.c_int]
        L.orc_krylov_band.argtypes = [c_i64, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_double, C.c_int,
                                      c_dp, c_dp, c_ip, c_dp]
        L.orc_krylov_csr_band.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.c_int, c_dp, C.c_int, C.c_int,
                                          C.c_double, C.c_int, c_dp, c_dp, c_ip, c_dp]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


# ------------------------------------------------------------------ generators
def gen_band(n: int, k: int, seed: int = 20140601, delta: float = 1.2) -> np.ndarray:
    a = np.empty((n, 2 * k + 1), dtype=np.float64)
    lib().orc_gen_band(n, k, seed, delta, _dp(a))
    return a


def gen_vec(n: int, seed: int = 20140601) -> np.ndarray:
    u = np.empty(n, dtype=np.float64)
    lib().orc_gen_vec(n, seed, _dp(u))
    return u


def u01(seed: int, counter: int) -> float:
    return lib().orc_u01(seed, counter)


# ------------------------------------------------------------------ reference restatements
def _csr(ia, ja, a):
    return (np.ascontiguousarray(ia, dtype=np.int32), np.ascontiguousarray(ja, dtype=np.int32),
            np.ascontiguousarray(a, dtype=np.float64))


def band_select(ia, ja, a, kmax: int, frac: float):
    """MatCreateSubMatrixBanded k/frac decision (src/matbanded.c:38-56,104-105)."""
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    k = C.c_int(0)
    f = C.c_double(0.0)
    rc = lib().orc_band_select(n, _ip(ia), _ip(ja), _dp(a), kmax, frac, C.byref(k), C.byref(f))
    if rc:
        raise ValueError("kmax > n is out of bounds in the reference (src/matbanded.c:54)")
    return k.value, f.value


def band_extract_csr(ia, ja, a, k: int):
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    ib = np.zeros(n + 1, dtype=np.int32)
    jb = np.zeros(len(ja), dtype=np.int32)
    b = np.zeros(len(ja), dtype=np.float64)
    nnz = lib().orc_band_extract_csr(n, _ip(ia), _ip(ja), _dp(a), k, _ip(ib), _ip(jb), _dp(b))
    return ib, jb[:nnz].copy(), b[:nnz].copy()


def csr_to_band(ia, ja, a, k: int) -> np.ndarray:
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    band = np.zeros((n, 2 * k + 1), dtype=np.float64)
    lib().orc_csr_to_band(n, _ip(ia), _ip(ja), _dp(a), k, _dp(band))
    return band


def mat_permute_csr(ia, ja, a, rowp, colp):
    """PETSc MatPermute as used at src/kspreorder.c:20: B(i,j) = A(rowp[i], colp[j])."""
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    rowp = np.ascontiguousarray(rowp, dtype=np.int32)
    colp = np.ascontiguousarray(colp, dtype=np.int32)
    ib = np.zeros(n + 1, dtype=np.int32)
    jb = np.zeros(len(ja), dtype=np.int32)
    b = np.zeros(len(ja), dtype=np.float64)
    rc = lib().orc_mat_permute_csr(n, _ip(ia), _ip(ja), _dp(a), _ip(rowp), _ip(colp), _ip(ib), _ip(jb), _dp(b))
    if rc:
        raise ValueError("not a permutation")
    return ib, jb, b


def vec_permute(x, idx, inverse: bool = False) -> np.ndarray:
    """PETSc VecPermute as used at src/kspreorder.c:122-127 (returns a permuted copy)."""
    y = np.array(x, dtype=np.float64, copy=True)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    lib().orc_vec_permute(len(y), _dp(y), _ip(idx), int(inverse))
    return y


def awbm(ia, ja, a):
    """MatGetOrdering_AWBM (src/petsc_mat_awbm.c:42-225): returns (permR, permC=identity, match)."""
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    perm = np.zeros(n, dtype=np.int32)
    match = np.zeros(n, dtype=np.int32)
    rc = lib().orc_awbm(n, _ip(ia), _ip(ja), _dp(a), _ip(perm), _ip(match))
    if rc:
        raise RuntimeError(f"AWBM failed ({rc})")
    return perm, np.arange(n, dtype=np.int32), match


def have_mc64() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libmc64ref.so"))


def wbm(ia, ja, a):
    """MatGetOrdering_WBM (src/petsc_mat_wbm.c:13-61) around the reference's own MC64 job 5.

    The wrapper hands the 1-based CSR arrays to MC64 as if they were CSC (:29,:52), with
    icntl = (0,0,0,0,4), cntl = 0 (:46-51), and returns row = identity, col = perm-1 (:55-58).
    Returns (row_is, col_is, num, dw) with dw the (discarded, :56) scaling workspace.
    The caller must pass a structurally symmetric pattern (MatGetRowIJ symmetric=TRUE hazard,
    SURVEY.md 8a-9) for the arrays to agree with the reference's.
    """
    global _MC64
    if _MC64 is None:
        build()
        _MC64 = C.CDLL(os.path.join(_HERE, "_ref", "libmc64ref.so"))
        _MC64.HSLmc64AD.restype = C.c_int
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    ia1 = (ia + 1).astype(np.int32)
    ja1 = (ja + 1).astype(np.int32)
    av = a.copy()
    nnz = int(ia[n])
    liw = 3 * n + 2 * n
    ldw = n + 2 * n + nnz
    iw = np.zeros(liw, dtype=np.int32)
    dw = np.zeros(ldw, dtype=np.float64)
    perm = np.zeros(n, dtype=np.int32)
    icntl = np.array([0, 0, 0, 0, 4], dtype=np.int32)
    cntl = np.array([0.0], dtype=np.float64)
    info = np.zeros(10, dtype=np.int32)
    job, m, nn, ne, num = C.c_int(5), C.c_int(n), C.c_int(n), C.c_int(nnz), C.c_int(0)
    cliw, cldw = C.c_int(liw), C.c_int(ldw)
    rc = _MC64.HSLmc64AD(C.byref(job), C.byref(m), C.byref(nn), C.byref(ne), _ip(ia1), _ip(ja1), _dp(av),
                         C.byref(num), _ip(perm), C.byref(cliw), _ip(iw), C.byref(cldw), _dp(dw),
                         _ip(icntl), _dp(cntl), _ip(info))
    if rc:
        raise RuntimeError(f"HSLmc64AD returned {rc}")
    return np.arange(n, dtype=np.int32), (perm - 1).astype(np.int32), num.value, dw


# ------------------------------------------------------------------ exact band solve
def band_lu(a: np.ndarray, boost: float = 0.0):
    """In-place no-pivot LU (rows layout); returns (lu, nboost)."""
    n, bw = a.shape
    lu = np.array(a, dtype=np.float64, order="C", copy=True)
    nb = lib().orc_band_lu(n, (bw - 1) // 2, _dp(lu), boost)
    return lu, nb


def band_solve(lu: np.ndarray, b: np.ndarray) -> np.ndarray:
    n, bw = lu.shape
    x = np.array(b, dtype=np.float64, order="C", copy=True)
    nrhs = 1 if x.ndim == 1 else x.shape[0]
    lib().orc_band_solve(n, (bw - 1) // 2, _dp(lu), _dp(x), nrhs, n)
    return x


def band_mult(a: np.ndarray, x: np.ndarray) -> np.ndarray:
    n, bw = a.shape
    y = np.empty(n, dtype=np.float64)
    xx = np.ascontiguousarray(x, dtype=np.float64)
    lib().orc_band_mult(n, (bw - 1) // 2, _dp(a), _dp(xx), _dp(y))
    return y


class Spike:
    """CPU truncated SPIKE (same algorithm as the GPU path), partition-parallel with OpenMP."""

    def __init__(self, n, k, nparts, align=8, tip_rows=0, boost=0.0):
        self.n, self.k, self.nparts = n, k, nparts
        self._h = lib().orc_spike_create(n, k, nparts, align, tip_rows, boost)
        self.lu = None

    def factor(self, a: np.ndarray, nthreads: int = 0, inplace: bool = False):
        self.lu = a if inplace else np.array(a, dtype=np.float64, order="C", copy=True)
        return lib().orc_spike_factor(self._h, _dp(self.lu), nthreads or lib().orc_num_threads())

    def solve(self, b: np.ndarray, nthreads: int = 0) -> np.ndarray:
        x = np.empty(self.n, dtype=np.float64)
        bb = np.ascontiguousarray(b, dtype=np.float64)
        lib().orc_spike_solve(self._h, _dp(self.lu), _dp(bb), _dp(x), nthreads or lib().orc_num_threads())
        return x

    def vb(self, i):
        return np.ctypeslib.as_array(lib().orc_spike_vb(self._h, i), shape=(self.k, self.k)).copy()

    def wt(self, i):
        return np.ctypeslib.as_array(lib().orc_spike_wt(self._h, i), shape=(self.k, self.k)).copy()

    def part_start(self, p):
        return lib().orc_spike_part_start(self._h, p)

    def __del__(self):
        try:
            lib().orc_spike_destroy(self._h)
        except Exception:
            pass


GMRES, BICGSTAB = 0, 1


def krylov_band(a, lu, b, method=GMRES, restart=30, rtol=1e-5, maxit=10000):
    n, bw = a.shape
    x = np.zeros(n)
    its = C.c_int(0)
    rn = C.c_double(0.0)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    rc = lib().orc_krylov_band(n, (bw - 1) // 2, _dp(a), _dp(lu), method, restart, rtol, maxit, _dp(bb), _dp(x),
                               C.byref(its), C.byref(rn))
    return x, its.value, rn.value, rc


def krylov_csr_band(ia, ja, a, lu, b, method=BICGSTAB, restart=30, rtol=1e-5, maxit=10000):
    ia, ja, a = _csr(ia, ja, a)
    n = len(ia) - 1
    k = (lu.shape[1] - 1) // 2
    x = np.zeros(n)
    its = C.c_int(0)
    rn = C.c_double(0.0)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    rc = lib().orc_krylov_csr_band(n, _ip(ia), _ip(ja), _dp(a), k, _dp(lu), method, restart, rtol, maxit,
                                   _dp(bb), _dp(x), C.byref(its), C.byref(rn))
    return x, its.value, rn.value, rc


def num_threads() -> int:
    return lib().orc_num_threads()


# ------------------------------------------------------------------ block LU (factor-entry parity)
def _gj_inverse_nopivot(D: np.ndarray, boost: float):
    """In-place Gauss-Jordan inverse without pivoting and with the same boosting rule as band_lu."""
    n = D.shape[0]
    A = D.copy()
    nb = 0
    for k in range(n):
        piv = A[k, k]
        if abs(piv) < boost:
            piv = -boost if piv < 0.0 else boost
            nb += 1
        rc = 1.0 / piv
        f = A[:, k] * rc
        f[k] = 0.0
        rowk = A[k, :].copy()
        A -= np.outer(f, rowk)
        A[k, :] = rowk * rc
        A[:, k] = -f
        A[k, k] = rc
    return A, nb


def block_lu(a: np.ndarray, tile: int = 8, boost: float = 0.0):
    """Block LU of the band (rows layout) grouped by `tile` pivots, the factorisation the GPU kernel
    stores: A~(I,J) below the diagonal, Ub(I,J) = D_I^-1 A~(I,J) above it, D_I^-1 in the diagonal tile.
    Same Schur complements / pivots / boosting as the scalar no-pivot LU (band_lu).  Returns the
    factors in a rows layout widened to the tile band (half-width kw = tile*ceil(k/tile) + tile-1)."""
    n, bw = a.shape
    k = (bw - 1) // 2
    kt = max(2, -(-k // tile))
    kw = tile * kt + tile - 1
    wide = np.zeros((n, 2 * kw + 1))
    wide[:, kw - k:kw + k + 1] = a
    nboost = 0
    nt = -(-n // tile)
    for s in range(nt):
        r0 = s * tile
        r1 = min(n, r0 + tile)
        e = min(n, r0 + tile * (kt + 1))           # rows/cols touched by this step
        idx = np.arange(r0, e)
        ii, jj = np.meshgrid(idx, idx, indexing="ij")
        dd = jj - ii + kw
        ok = (dd >= 0) & (dd <= 2 * kw)
        Wd = np.zeros((e - r0, e - r0))
        Wd[ok] = wide[ii[ok], dd[ok]]
        m = r1 - r0
        Dinv, nb = _gj_inverse_nopivot(Wd[:m, :m], boost)
        nboost += nb
        Ub = Dinv @ Wd[:m, m:]
        Wd[m:, m:] -= Wd[m:, :m] @ Ub
        Wd[:m, m:] = Ub
        Wd[:m, :m] = Dinv
        wide[ii[ok], dd[ok]] = Wd[ok]
    return wide, nboost, kw
