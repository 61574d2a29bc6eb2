"""On-disk formats of the reference's drivers (spike_petsc_b200/host/matio.c): PETSc binary Mat/Vec as MatLoad reads
them (src/testbed2.c:93-96) and the MatrixMarket export of src/wbm.c:520-523.  CPU only."""
import ctypes as C
import os
import struct

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "spike_petsc_b200", "lib", "libspike_petsc.so")


@pytest.fixture(scope="module")
def io():
    assert os.path.exists(LIB), "libspike_petsc.so missing: run make"
    L = C.CDLL(LIB)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.SpkMatLoadBinary.argtypes = [C.c_char_p, ip, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp)]
    L.SpkMatWriteBinary.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.SpkVecLoadBinary.argtypes = [C.c_char_p, ip, C.POINTER(dp)]
    L.SpkVecWriteBinary.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    L.SpkMatWriteMatrixMarket.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.SpkFree.argtypes = [C.c_void_p]
    return L


def _load_mat(L, path):
    m, n = C.c_int(), C.c_int()
    ia, ja, a = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    rc = L.SpkMatLoadBinary(path.encode(), C.byref(m), C.byref(n), C.byref(ia), C.byref(ja), C.byref(a))
    if rc:
        return rc, None
    nz = ia[m.value]
    A = sp.csr_matrix((np.ctypeslib.as_array(a, (max(nz, 1),))[:nz].copy(), np.ctypeslib.as_array(ja, (max(nz, 1),))[:nz].copy(),
                       np.ctypeslib.as_array(ia, (m.value + 1,)).copy()), shape=(m.value, n.value))
    for p in (ia, ja, a):
        L.SpkFree(p)
    return 0, A


# the 3x3 matrix "From HC64 documentation" of src/wbm.c:483-497
WBM3 = sp.csr_matrix(np.array([[0.0, 8.0, 3.0], [0.0, 2.0, 1.0], [4.0, 0.0, 0.0]]))


def test_petsc_binary_known_bytes(io, tmp_path):
    """Byte-for-byte what PETSc's MatView(binary) writes for the 3x3 matrix: big-endian int32/float64."""
    p = str(tmp_path / "wbm3.bin")
    A = WBM3
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    assert io.SpkMatWriteBinary(p.encode(), 3, 3, ia.ctypes.data, ja.ctypes.data, a.ctypes.data) == 0
    expect = struct.pack(">4i", 1211216, 3, 3, 5) + struct.pack(">3i", 2, 2, 1) + struct.pack(">5i", 1, 2, 1, 2, 0) \
        + struct.pack(">5d", 8.0, 3.0, 2.0, 1.0, 4.0)
    assert open(p, "rb").read() == expect
    rc, B = _load_mat(io, p)
    assert rc == 0 and (B != A).nnz == 0 and np.array_equal(B.indptr, A.indptr) and np.array_equal(B.indices, A.indices)


def test_petsc_binary_round_trip_random(io, tmp_path):
    rng = np.random.default_rng(5)
    A = sp.random(500, 500, density=0.02, random_state=rng, format="csr") + sp.eye(500, format="csr") * 3.0
    A = A.tocsr(); A.sort_indices()
    p = str(tmp_path / "a.bin")
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    assert io.SpkMatWriteBinary(p.encode(), 500, 500, ia.ctypes.data, ja.ctypes.data, a.ctypes.data) == 0
    rc, B = _load_mat(io, p)
    assert rc == 0
    assert np.array_equal(B.indptr, ia) and np.array_equal(B.indices, ja) and np.array_equal(B.data, a)   # bit exact
    x = rng.standard_normal(500)
    pv = str(tmp_path / "x.bin")
    assert io.SpkVecWriteBinary(pv.encode(), 500, x.ctypes.data) == 0
    assert open(pv, "rb").read()[:8] == struct.pack(">2i", 1211214, 500)
    n, v = C.c_int(), C.POINTER(C.c_double)()
    assert io.SpkVecLoadBinary(pv.encode(), C.byref(n), C.byref(v)) == 0
    assert n.value == 500 and np.array_equal(np.ctypeslib.as_array(v, (500,)), x)
    io.SpkFree(v)


def test_petsc_binary_error_paths(io, tmp_path):
    assert _load_mat(io, str(tmp_path / "missing.bin"))[0] == 65          # PETSC_ERR_FILE_OPEN
    p = str(tmp_path / "vec_as_mat.bin")
    open(p, "wb").write(struct.pack(">2i", 1211214, 2) + struct.pack(">2d", 1.0, 2.0))
    assert _load_mat(io, p)[0] in (66, 79)                                # wrong class id / short file
    p2 = str(tmp_path / "trunc.bin")
    open(p2, "wb").write(struct.pack(">4i", 1211216, 3, 3, 5) + struct.pack(">3i", 2, 2, 1))
    assert _load_mat(io, p2)[0] == 66                                     # PETSC_ERR_FILE_READ
    p3 = str(tmp_path / "badlen.bin")
    open(p3, "wb").write(struct.pack(">4i", 1211216, 2, 2, 2) + struct.pack(">2i", 2, 1) + struct.pack(">2i", 0, 1) + struct.pack(">2d", 1.0, 2.0))
    assert _load_mat(io, p3)[0] == 79                                     # row lengths do not add up to nnz
    p4 = str(tmp_path / "badcol.bin")
    open(p4, "wb").write(struct.pack(">4i", 1211216, 2, 2, 2) + struct.pack(">2i", 1, 1) + struct.pack(">2i", 0, 7) + struct.pack(">2d", 1.0, 2.0))
    assert _load_mat(io, p4)[0] == 79                                     # column index out of range


def test_matrixmarket_export(io, tmp_path):
    A = WBM3
    ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    p = str(tmp_path / "wbm3.mtx")
    assert io.SpkMatWriteMatrixMarket(p.encode(), 3, 3, ia.ctypes.data, ja.ctypes.data, a.ctypes.data, 0) == 0
    assert open(p).read() == "%%MatrixMarket matrix coordinate real general\n3 3 5\n1 2 8\n1 3 3\n2 2 2\n2 3 1\n3 1 4\n"
    rng = np.random.default_rng(9)
    B = sp.random(60, 40, density=0.1, random_state=rng, format="csr"); B.sort_indices()
    ib, jb, b = B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data.astype(np.float64)
    p2 = str(tmp_path / "b.mtx")
    assert io.SpkMatWriteMatrixMarket(p2.encode(), 60, 40, ib.ctypes.data, jb.ctypes.data, b.ctypes.data, 17) == 0
    R = scipy.io.mmread(p2).tocsr()
    assert R.shape == (60, 40) and abs(R - B).max() == 0.0               # %.17g round trips exactly
