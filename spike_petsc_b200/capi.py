"""ctypes binding of include/spike_b200.h (the C-ABI drop-in boundary).

Pointers cross the boundary as plain addresses: numpy arrays for host memory, or raw device
addresses (e.g. ``torch.Tensor.data_ptr()``) when the context was created with mem=MEM_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GMRES, BCGS = 0, 1
LAYOUT_ROWS, LAYOUT_DIAGS = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
BND_WT_FIRST, BND_REMOTE_WT, BND_G_TOP, BND_REMOTE_G_TOP, BND_X_BOT, BND_REMOTE_X_BOT, BND_HALO_LEFT, BND_HALO_RIGHT = 0, 2, 3, 4, 5, 6, 7, 8


class SpikeError(RuntimeError):
    pass


class Opts(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("partitions", C.c_int), ("tip_tiles", C.c_int),
                ("boost_rel", C.c_double), ("mem", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("row_offset", C.c_int64), ("n_global", C.c_int64)]


class Info(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_padded", C.c_int64), ("k", C.c_int), ("k_padded", C.c_int), ("kt", C.c_int),
                ("partitions", C.c_int), ("tip_tiles", C.c_int), ("boosted_pivots", C.c_int64), ("factored", C.c_int),
                ("frac", C.c_double), ("anorm_max", C.c_double), ("factor_ms", C.c_double), ("solve_ms", C.c_double),
                ("band_bytes", C.c_int64), ("kernel_launches", C.c_int), ("stage_ms", C.c_double * 8)]


def library_path() -> str:
    # SPIKE_B200_LIB: an alternative build of the same library (kernel experiments of tools/, never set in tests/bench)
    return os.environ.get("SPIKE_B200_LIB") or os.path.join(_HERE, "lib", "libspike_b200.so")


def exported_symbols() -> list[str]:
    """Names declared in include/spike_b200.h (used by the CPU-side load test)."""
    hdr = os.path.join(os.path.dirname(_HERE), "include", "spike_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(spk_[a-z_0-9]+)\s*\(", txt)))


def lib():
    """Load libspike_b200.so; fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise SpikeError(f"{path} is missing: run `make` (or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(path)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.spk_default_opts.argtypes = [C.POINTER(Opts)]
        L.spk_last_error.restype = C.c_char_p
        L.spk_last_error.argtypes = [vp]
        L.spk_version.restype = C.c_char_p
        L.spk_create.argtypes = [C.POINTER(vp), C.POINTER(Opts)]
        L.spk_destroy.argtypes = [C.POINTER(vp)]
        L.spk_set_band_dense.argtypes = [vp, C.c_int64, C.c_int, vp, C.c_int, C.c_int]
        L.spk_set_band_csr.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, ip, dp]
        L.spk_set_band_synthetic.argtypes = [vp, C.c_int64, C.c_int, C.c_uint64, C.c_double]
        L.spk_get_band_rows.argtypes = [vp, vp]
        L.spk_factor.argtypes = [vp]
        L.spk_solve.argtypes = [vp, vp, vp, C.c_int]
        L.spk_mult.argtypes = [vp, vp, vp]
        L.spk_keep_original.argtypes = [vp, C.c_int]
        L.spk_set_scaling.argtypes = [vp, vp, vp]
        L.spk_permute.argtypes = [vp, vp, C.c_int, vp, C.c_int64]
        L.spk_set_operator_csr.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.spk_krylov.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, ip, dp, ip]
        L.spk_view.argtypes = [vp, C.POINTER(Info)]
        L.spk_set_timing.argtypes = [vp, C.c_int]
        L.spk_check.argtypes = [vp, dp]
        L.spk_awbm_csr.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        L.spk_tip_size.argtypes = [vp, ip]
        L.spk_get_boundary.argtypes = [vp, C.c_int, vp]
        L.spk_set_boundary.argtypes = [vp, C.c_int, vp]
        L.spk_factor_phase.argtypes = [vp, C.c_int]
        L.spk_solve_phase.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.spk_reserve_rhs.argtypes = [vp, C.c_int]
        L.spk_peer_mailbox_create.argtypes = [vp, vp, C.POINTER(vp)]
        L.spk_peer_mailbox_attach.argtypes = [vp, C.c_int, vp, vp]
        L.spk_peer_post.argtypes = [vp, C.c_int]
        L.spk_peer_wait.argtypes = [vp, C.c_int]
        L.spk_peer_check.argtypes = [vp]
        _LIB = L
    return _LIB


def _addr(a):
    """numpy array -> address; int -> itself (device pointer); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Spike:
    """Python mirror of one spk_ctx.  Mirrors the PC_Banded life cycle of the reference
    (src/matbanded.c:111-283): create -> set band (k,frac in/out) -> factor (PCSetUp) -> solve (PCApply)."""

    def __init__(self, device=0, partitions=0, tip_tiles=0, boost_rel=1e-13, mem=MEM_HOST, stream=None,
                 rank=0, nranks=1, row_offset=0, n_global=0, timing=True):
        L = lib()
        o = Opts()
        L.spk_default_opts(C.byref(o))
        o.device, o.partitions, o.tip_tiles, o.boost_rel, o.mem = device, partitions, tip_tiles, boost_rel, mem
        o.stream = stream
        o.rank, o.nranks, o.row_offset, o.n_global = rank, nranks, row_offset, n_global
        self._h = C.c_void_p()
        self.mem = mem
        rc = L.spk_create(C.byref(self._h), C.byref(o))
        if rc:
            raise SpikeError(f"spk_create failed ({rc}): {L.spk_last_error(None).decode()}")
        self.n = 0
        self.k = 0
        # the library's default is no event timers (they cost ~2 us each between kernels); tests and tools want the
        # stage times of view(), bench.py switches them off for its timed region
        if timing:
            self.set_timing(True)

    def set_timing(self, level):
        """True / 2: all event timers; 1: only the band-LU stage; False / 0: none (the library default)."""
        level = 2 if level is True else (0 if level is False else int(level))
        self._ck(lib().spk_set_timing(self._h, level), "spk_set_timing")

    def _ck(self, rc, what):
        if rc:
            raise SpikeError(f"{what} failed ({rc}): {lib().spk_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().spk_destroy(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- band definition
    def set_band_dense(self, band: np.ndarray, k: int, layout=LAYOUT_ROWS):
        band = np.ascontiguousarray(band, dtype=np.float64)
        n = band.shape[0] if layout == LAYOUT_ROWS else band.shape[1]
        self._keep = band
        self._ck(lib().spk_set_band_dense(self._h, n, k, band.ctypes.data, layout, MEM_HOST), "spk_set_band_dense")
        self.n, self.k = n, k

    def set_band_dense_device(self, ptr: int, n: int, k: int, layout=LAYOUT_ROWS):
        self._ck(lib().spk_set_band_dense(self._h, n, k, ptr, layout, MEM_DEVICE), "spk_set_band_dense")
        self.n, self.k = n, k

    def set_band_synthetic(self, n: int, k: int, seed=20140601, delta=1.2):
        self._ck(lib().spk_set_band_synthetic(self._h, n, k, seed, delta), "spk_set_band_synthetic")
        self.n, self.k = n, k

    def set_band_csr(self, ia, ja, a, kmax: int, frac: float, rowperm=None, colperm=None):
        """MatCreateSubMatrixBanded(+MatPermute): returns the (k, frac) the reference would report."""
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ja = np.ascontiguousarray(ja, dtype=np.int32)
        a = np.ascontiguousarray(a, dtype=np.float64)
        rp = None if rowperm is None else np.ascontiguousarray(rowperm, dtype=np.int32)
        cp = None if colperm is None else np.ascontiguousarray(colperm, dtype=np.int32)
        n = len(ia) - 1
        k = C.c_int(kmax)
        f = C.c_double(frac)
        self._ck(lib().spk_set_band_csr(self._h, n, _addr(ia), _addr(ja), _addr(a), _addr(rp), _addr(cp),
                                        C.byref(k), C.byref(f)), "spk_set_band_csr")
        self.n, self.k = n, k.value
        return k.value, f.value

    def set_operator_csr(self, ia, ja, a, rowperm=None, colperm=None):
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ja = np.ascontiguousarray(ja, dtype=np.int32)
        a = np.ascontiguousarray(a, dtype=np.float64)
        rp = None if rowperm is None else np.ascontiguousarray(rowperm, dtype=np.int32)
        cp = None if colperm is None else np.ascontiguousarray(colperm, dtype=np.int32)
        self._ck(lib().spk_set_operator_csr(self._h, len(ia) - 1, _addr(ia), _addr(ja), _addr(a), _addr(rp), _addr(cp)),
                 "spk_set_operator_csr")

    def set_scaling(self, rscale, cscale):
        """band <- diag(r) band diag(c) before factor(); solve() still returns the solution of the original system."""
        if self.mem == MEM_HOST:
            rscale = np.ascontiguousarray(rscale, dtype=np.float64)
            cscale = np.ascontiguousarray(cscale, dtype=np.float64)
        self._ck(lib().spk_set_scaling(self._h, _addr(rscale), _addr(cscale)), "spk_set_scaling")

    def keep_original(self, keep=True):
        self._ck(lib().spk_keep_original(self._h, int(keep)), "spk_keep_original")

    def get_band_rows(self) -> np.ndarray:
        out = np.empty((self.n, 2 * self.k + 1), dtype=np.float64)
        self._ck(lib().spk_get_band_rows(self._h, out.ctypes.data), "spk_get_band_rows")
        return out

    # ---- hot path
    def factor(self):
        self._ck(lib().spk_factor(self._h), "spk_factor")

    def solve(self, b, x=None, nrhs=1):
        """Host mode: numpy in, numpy out.  Device mode: b/x are device addresses (x may equal b)."""
        if self.mem == MEM_HOST:
            b = np.ascontiguousarray(b, dtype=np.float64)
            out = np.empty_like(b)
            self._ck(lib().spk_solve(self._h, b.ctypes.data, out.ctypes.data, nrhs), "spk_solve")
            return out
        self._ck(lib().spk_solve(self._h, _addr(b), _addr(x if x is not None else b), nrhs), "spk_solve")
        return x if x is not None else b

    def mult(self, x, y=None):
        if self.mem == MEM_HOST:
            x = np.ascontiguousarray(x, dtype=np.float64)
            out = np.empty_like(x)
            self._ck(lib().spk_mult(self._h, x.ctypes.data, out.ctypes.data), "spk_mult")
            return out
        self._ck(lib().spk_mult(self._h, _addr(x), _addr(y)), "spk_mult")
        return y

    def permute(self, idx, v, inverse=False):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if self.mem == MEM_HOST:
            v = np.array(v, dtype=np.float64, copy=True)
            self._ck(lib().spk_permute(self._h, idx.ctypes.data, int(inverse), v.ctypes.data, len(v)), "spk_permute")
            return v
        self._ck(lib().spk_permute(self._h, idx.ctypes.data, int(inverse), _addr(v), len(idx)), "spk_permute")
        return v

    def krylov(self, b, method=GMRES, restart=30, rtol=1e-5, maxit=10000, x=None):
        its, conv, rn = C.c_int(0), C.c_int(0), C.c_double(0.0)
        if self.mem == MEM_HOST:
            b = np.ascontiguousarray(b, dtype=np.float64)
            out = np.zeros_like(b)
            self._ck(lib().spk_krylov(self._h, method, restart, rtol, maxit, b.ctypes.data, out.ctypes.data,
                                      C.byref(its), C.byref(rn), C.byref(conv)), "spk_krylov")
            return out, its.value, rn.value, bool(conv.value)
        self._ck(lib().spk_krylov(self._h, method, restart, rtol, maxit, _addr(b), _addr(x), C.byref(its),
                                  C.byref(rn), C.byref(conv)), "spk_krylov")
        return x, its.value, rn.value, bool(conv.value)

    # ---- sharded (multi-GPU) split-phase interface; device memory only
    def tip_size(self) -> int:
        kp = C.c_int(0)
        self._ck(lib().spk_tip_size(self._h, C.byref(kp)), "spk_tip_size")
        return kp.value

    overlapped_factor = True   # supports factor phases 10/11 (W^(t) exchange overlapped with the band LU)

    def factor_phase(self, phase: int):
        self._ck(lib().spk_factor_phase(self._h, phase), f"spk_factor_phase({phase})")

    def solve_phase(self, phase: int, b=None, x=None, nrhs: int = 1):
        self._ck(lib().spk_solve_phase(self._h, phase, _addr(b), _addr(x), nrhs), f"spk_solve_phase({phase})")

    def reserve_rhs(self, nrhs: int):
        """Size the boundary exchange buffers for nrhs columns per sharded solve (before peer_create)."""
        self._ck(lib().spk_reserve_rhs(self._h, nrhs), "spk_reserve_rhs")

    def get_boundary(self, which: int, buf):
        self._ck(lib().spk_get_boundary(self._h, which, _addr(buf)), f"spk_get_boundary({which})")

    def set_boundary(self, which: int, buf):
        self._ck(lib().spk_set_boundary(self._h, which, _addr(buf)), f"spk_set_boundary({which})")

    # ---- the same exchanges through NVLink peer memory (csrc/peer.cu)
    peer_capable = True

    def peer_create(self):
        """-> (64-byte CUDA IPC handle, device address) of this rank's mailbox."""
        h = C.create_string_buffer(64)
        p = C.c_void_p(0)
        self._ck(lib().spk_peer_mailbox_create(self._h, h, C.byref(p)), "spk_peer_mailbox_create")
        return bytes(h.raw), p.value

    def peer_attach(self, side: int, handle: bytes = None, ptr: int = None):
        hb = C.create_string_buffer(handle, 64) if handle is not None else None
        self._ck(lib().spk_peer_mailbox_attach(self._h, side, hb, C.c_void_p(ptr) if ptr else None), f"spk_peer_mailbox_attach({side})")

    def peer_post(self, which: int):
        self._ck(lib().spk_peer_post(self._h, which), f"spk_peer_post({which})")

    def peer_wait(self, which: int):
        self._ck(lib().spk_peer_wait(self._h, which), f"spk_peer_wait({which})")

    def peer_check(self):
        self._ck(lib().spk_peer_check(self._h), "spk_peer_check")

    def awbm(self, ia, ja, a, want_scalings=False):
        """MatGetOrdering_AWBM on the GPU (spk_awbm_csr) -> (permR, match, stats[, scalR, scalC])."""
        ia = np.ascontiguousarray(ia, dtype=np.int32)
        ja = np.ascontiguousarray(ja, dtype=np.int32)
        a = np.ascontiguousarray(a, dtype=np.float64)
        n = ia.size - 1
        permR, match, stats = np.empty(n, np.int32), np.empty(n, np.int32), np.zeros(4, np.int32)
        sr = np.empty(n) if want_scalings else None
        sc = np.empty(n) if want_scalings else None
        self._ck(lib().spk_awbm_csr(self._h, n, _addr(ia), _addr(ja), _addr(a), _addr(permR), _addr(match), _addr(sr), _addr(sc),
                                    _addr(stats)), "spk_awbm_csr")
        return (permR, match, stats, sr, sc) if want_scalings else (permR, match, stats)

    def check(self) -> float:
        """||x - v|| / ||v|| of one probe solve through the kept unfactored band (spk_check)."""
        e = C.c_double(0.0)
        self._ck(lib().spk_check(self._h, C.byref(e)), "spk_check")
        return e.value

    def view(self) -> dict:
        info = Info()
        self._ck(lib().spk_view(self._h, C.byref(info)), "spk_view")
        d = {f[0]: getattr(info, f[0]) for f in Info._fields_}
        d["stage_ms"] = list(info.stage_ms)
        return d
