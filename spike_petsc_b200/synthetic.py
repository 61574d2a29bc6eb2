"""Counter-based synthetic inputs of SURVEY.md 8d that are built on the host: the sparse nonsymmetric matrix of
BASELINE config 4 (C4) and the two-stage "wbm + fiedler" ordering flow of the reference's testbed driver.

The banded inputs (C1, C2, C3, C5) are generated directly in device memory by spk_set_band_synthetic; C4 is a CSR
matrix, the form the reference's drivers load from disk (src/testbed2.c:93-96), so it is assembled here with numpy.
Every number comes from splitmix64(seed ^ counter), the generator the CPU oracle and the CUDA kernels share.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def u01(seed: int, counter: np.ndarray) -> np.ndarray:
    """uniform [0,1): (splitmix64(seed ^ counter) >> 11) * 2^-53  (oracle/spike_oracle.c orc_u01, csrc/common.cuh)."""
    c = np.asarray(counter, dtype=np.uint64)
    return (splitmix64(np.uint64(seed) ^ c) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hashed_permutation(n: int, seed: int) -> np.ndarray:
    """A fixed pseudo-random permutation: stable argsort of the hashed indices."""
    return np.argsort(splitmix64(np.uint64(seed) ^ np.arange(n, dtype=np.uint64)), kind="stable").astype(np.int64)


def hashed_involution(n: int, seed: int) -> np.ndarray:
    """A fixed pseudo-random permutation that is its own inverse (disjoint transpositions of a hashed pairing).
    The reference exposes MC64's ROW matching as the COLUMN index set (src/petsc_mat_wbm.c:55-58); the two
    coincide exactly when the scrambling row permutation is an involution, which keeps config 4 meaningful under the
    reference's own semantics."""
    p = hashed_permutation(n, seed)
    r = np.arange(n, dtype=np.int64)
    m = (n // 2) * 2
    a, b = p[0:m:2], p[1:m:2]
    r[a], r[b] = b, a
    return r


def c4_matrix(n: int = 2_000_000, half_bandwidth: int = 2000, pairs_per_row: int = 5, delta: float = 1.2,
              seeds=(20140602, 20140603, 20140604), scramble: bool = True):
    """structurally symmetric random band pattern (pairs_per_row off-diagonal pairs per row inside
    +-half_bandwidth), nonsymmetric values uniform(-1,1), diagonal = delta * sum|off-diagonal| of its row; then a fixed
    symmetric permutation Q (destroys the band) and a fixed row permutation R (an involution, moves the dominant entries
    off the diagonal).  Returns (A' as scipy CSR with sorted indices, Q, R); A' = (A0[Q][:, Q])[R, :]."""
    s_pat, s_q, s_r = seeds
    i = np.repeat(np.arange(n, dtype=np.int64), pairs_per_row)
    p = np.tile(np.arange(pairs_per_row, dtype=np.int64), n)
    d = 1 + np.floor(u01(s_pat, (i * 8 + p).astype(np.uint64)) * half_bandwidth).astype(np.int64)
    j = i + d
    ok = j < n
    i, j = i[ok], j[ok]
    rows = np.concatenate([i, j])
    cols = np.concatenate([j, i])
    key = np.unique(rows * n + cols)                       # the pattern as a set: duplicates collapse
    rows, cols = key // n, key % n
    vals = 2.0 * u01(s_pat ^ 0x5A5A5A5A, key.astype(np.uint64)) - 1.0   # value hashed from (row, col): nonsymmetric
    rowsum = np.bincount(rows, weights=np.abs(vals), minlength=n)
    diag = delta * rowsum
    diag[diag == 0.0] = 1.0
    A0 = sp.csr_matrix((np.concatenate([vals, diag]), (np.concatenate([rows, np.arange(n)]), np.concatenate([cols, np.arange(n)]))),
                       shape=(n, n))
    A0.sort_indices()
    if not scramble:
        return A0, np.arange(n), np.arange(n)
    Q = hashed_permutation(n, s_q)
    R = hashed_involution(n, s_r)
    A1 = A0[Q][:, Q]
    A2 = sp.csr_matrix(A1[R])
    A2.sort_indices()
    return A2, Q, R


def compose_wbm_then_symmetric(col1: np.ndarray, sym2: np.ndarray):
    """Two-stage ordering of the reference's testbed driver (-mat_ordering_type wbm -mat_ordering_type2 fiedler,
    src/testbed.c:200-284): stage 1 permutes columns by col1 (rows stay), stage 2 permutes rows and columns of the
    result by sym2.  As ONE MatPermute: B(i,j) = A(row[i], col[j]) with row = sym2, col = col1[sym2]."""
    col1 = np.asarray(col1, dtype=np.int64)
    sym2 = np.asarray(sym2, dtype=np.int64)
    return sym2.astype(np.int32), col1[sym2].astype(np.int32)
