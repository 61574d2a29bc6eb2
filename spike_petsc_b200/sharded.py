"""Row-block sharding of one banded system across the GPUs of a box (one process per GPU).

SPIKE partitions shard naturally: rank r owns a contiguous block of rows (a multiple of 8) and runs
the same factor / solve kernels on it.  Only the data of the partition interface that straddles two
ranks crosses NVLink: one kp x kp spike tip W^(t) per boundary at factor time and two kp-vectors per
boundary per solve (kp = 8*ceil(K/8)); this replaces the PETSc MPI scatter a distributed Mat/Vec
would do on that path.  The exchanges are point-to-point `torch.distributed` operations, i.e.
ncclSend / ncclRecv over NVLink 5 when the backend is NCCL -- or, when the engine supports it (the CUDA
engine does: csrc/peer.cu), stores into the neighbour's mailbox in NVLink peer memory issued by the engine's own
kernels, with no host round per exchange (the NCCL path stays as the fallback when CUDA IPC is unavailable and
for SPIKE_B200_PEER=0).  The engine argument exists so the exchange protocol can be exercised on CPU (gloo)
with a reference engine (tests/test_sharded_cpu.py).
"""
from __future__ import annotations

import os
import sys

from . import capi


def shard_rows(n_global: int, world: int, k: int = 0):
    """Contiguous row blocks, every boundary a multiple of 8 rows (tile size) -- of 64 rows (super-block size) for
    the wide-band path, half-bandwidth k > 128."""
    al = 64 if k > 128 else 8
    tiles = (n_global + al - 1) // al
    bounds = [min(n_global, (tiles * r // world) * al) for r in range(world + 1)]
    bounds[-1] = n_global
    return bounds


class ShardedSpike:
    def __init__(self, engine, rank: int, world: int, dist=None, nrhs: int = 1):
        """engine: an object with the split-phase API of capi.Spike (factor_phase, solve_phase,
        get_boundary, set_boundary, tip_size) created for this rank's row block.
        nrhs: the most right-hand-side columns one solve() will carry (sizes the exchange buffers / mailboxes)."""
        self.e, self.rank, self.world, self.dist = engine, rank, world, dist
        self.nrhs = nrhs
        if nrhs > 1:
            engine.reserve_rhs(nrhs)
        self._bufs = None
        self._peer = None        # None: not decided yet; True: NVLink mailboxes; False: NCCL send/recv
        self._peer_verified = False

    # ---- NVLink peer mailboxes: every rank creates one, the CUDA IPC handles travel once over the process group
    def _peer_setup(self):
        if self._peer is not None:
            return self._peer
        self._peer = False
        if self.world == 1 or not getattr(self.e, "peer_capable", False) or os.environ.get("SPIKE_B200_PEER", "1") == "0":
            return False
        td = self.dist
        if td is None:
            import torch.distributed as td
        ok, handle = 1, b""
        try:
            handle, _ = self.e.peer_create()
        except Exception as exc:   # noqa: BLE001 -- any failure means "use NCCL", decided collectively below
            ok = 0
            print(f"[spike_b200] rank {self.rank}: no peer mailbox ({exc}); NCCL exchange", file=sys.stderr)
        got = [None] * self.world
        td.all_gather_object(got, (ok, handle))
        if all(g[0] for g in got):
            try:
                if self.rank > 0:
                    self.e.peer_attach(0, handle=got[self.rank - 1][1])
                if self.rank + 1 < self.world:
                    self.e.peer_attach(1, handle=got[self.rank + 1][1])
            except Exception as exc:   # noqa: BLE001
                ok = 0
                print(f"[spike_b200] rank {self.rank}: cannot map the neighbour's mailbox ({exc}); NCCL exchange", file=sys.stderr)
        else:
            ok = 0
        agreed = [None] * self.world
        td.all_gather_object(agreed, ok)
        self._peer = all(agreed)
        return self._peer

    def attach_local_peers(self, left_ptr=None, right_ptr=None):
        """Neighbour shards living in this process (several shards on one GPU): use their mailboxes directly."""
        self.e.peer_create()
        if left_ptr:
            self.e.peer_attach(0, ptr=left_ptr)
        if right_ptr:
            self.e.peer_attach(1, ptr=right_ptr)
        self._peer = True

    def check(self):
        """Synchronise and raise if a mailbox spin timed out (peer protocol only)."""
        if self._peer:
            self.e.peer_check()

    # ---- neighbour exchange: send `out` to rank+dir_, receive the matching buffer from rank-dir_
    def _shift(self, send_buf, recv_buf, direction: int):
        """direction=-1: send to the left neighbour / receive from the right one; +1 the reverse."""
        if self.world == 1:
            return
        import torch.distributed as td
        ops = []
        dst = self.rank + direction
        src = self.rank - direction
        if 0 <= dst < self.world:
            ops.append(td.P2POp(td.isend, send_buf, dst))
        if 0 <= src < self.world:
            ops.append(td.P2POp(td.irecv, recv_buf, src))
        if ops:
            for req in td.batch_isend_irecv(ops):
                req.wait()

    def _shift_start(self, send_buf, recv_buf, direction: int):
        """Like _shift, but returns the pending requests instead of waiting for them."""
        import torch.distributed as td
        ops = []
        dst = self.rank + direction
        src = self.rank - direction
        if 0 <= dst < self.world:
            ops.append(td.P2POp(td.isend, send_buf, dst))
        if 0 <= src < self.world:
            ops.append(td.P2POp(td.irecv, recv_buf, src))
        return td.batch_isend_irecv(ops) if ops else []

    def _alloc(self, like):
        import torch
        kp = self.e.tip_size()
        mk = lambda m: torch.zeros(m, dtype=torch.float64, device=like.device)  # noqa: E731
        self._bufs = {"wt_out": mk(kp * kp), "wt_in": mk(kp * kp), "v_out": mk(kp * self.nrhs), "v_in": mk(kp * self.nrhs)}
        self.kp = kp

    def _ptr(self, t):
        return self.e.buffer_address(t) if hasattr(self.e, "buffer_address") else t.data_ptr()

    def factor(self, like):
        """like: any tensor on this rank's device (used to allocate the exchange buffers)."""
        if self._bufs is None:
            self._alloc(like)
        b = self._bufs
        has_left, has_right = self.rank > 0, self.rank + 1 < self.world
        if self.world > 1 and self._peer_setup():
            # NVLink mailboxes: W^(t) of my first partition is stored into the left neighbour's memory by a kernel
            # queued before the band LU; the kernel that picks up the right neighbour's W^(t) is queued after it
            self.e.factor_phase(10)
            if has_left:
                self.e.peer_post(capi.BND_WT_FIRST)
            self.e.factor_phase(11)
            if has_right:
                self.e.peer_wait(capi.BND_REMOTE_WT)
            self.e.factor_phase(1)
            if has_right:
                self.e.factor_phase(2)
            return
        if self.world == 1 or not getattr(self.e, "overlapped_factor", False):
            self.e.factor_phase(0)
            self.e.factor_phase(1)
            if self.world > 1:
                if has_left:
                    self.e.get_boundary(capi.BND_WT_FIRST, self._ptr(b["wt_out"]))
                self._sync(like)
                self._shift(b["wt_out"], b["wt_in"], -1)
                if has_right:
                    self.e.set_boundary(capi.BND_REMOTE_WT, self._ptr(b["wt_in"]))
                    self.e.factor_phase(2)
            return
        # overlapped protocol: W^(t) of my first partition only needs the tip windows, so it travels to the
        # left neighbour on the NCCL stream while the band LU runs on the compute stream
        self.e.factor_phase(10)
        if has_left:
            self.e.get_boundary(capi.BND_WT_FIRST, self._ptr(b["wt_out"]))
        reqs = self._shift_start(b["wt_out"], b["wt_in"], -1)
        self.e.factor_phase(11)
        for req in reqs:
            req.wait()
        if has_right:
            self.e.set_boundary(capi.BND_REMOTE_WT, self._ptr(b["wt_in"]))
        self.e.factor_phase(1)
        if has_right:
            self.e.factor_phase(2)

    def solve(self, bvec, xvec, nrhs: int = 1):
        """bvec, xvec: this rank's rows of b and x (tensors on the rank's device; may alias); nrhs > 1: nrhs columns,
        column r at offset r * (local rows)."""
        if nrhs > self.nrhs:
            raise ValueError(f"solve with {nrhs} columns on a ShardedSpike created for {self.nrhs}")
        kw = {"nrhs": nrhs} if nrhs > 1 else {}
        if self._bufs is None:
            self._alloc(bvec)
        b = self._bufs
        has_left, has_right = self.rank > 0, self.rank + 1 < self.world
        if self.world > 1 and self._peer_setup():
            self.e.solve_phase(0, self._ptr(bvec), self._ptr(xvec), **kw)
            if has_left:
                self.e.peer_post(capi.BND_G_TOP)
            if has_right:
                self.e.peer_wait(capi.BND_REMOTE_G_TOP)
            self.e.solve_phase(1)
            if has_right:
                self.e.peer_post(capi.BND_X_BOT)
            if has_left:
                self.e.peer_wait(capi.BND_REMOTE_X_BOT)
            self.e.solve_phase(2)
            if not self._peer_verified:   # the first exchange proves the mapping works; later checks are the caller's
                self.e.peer_check()
                self._peer_verified = True
            return xvec
        self.e.solve_phase(0, self._ptr(bvec), self._ptr(xvec), **kw)
        if self.world > 1:
            if has_left:
                self.e.get_boundary(capi.BND_G_TOP, self._ptr(b["v_out"]))
            self._sync(bvec)
            self._shift(b["v_out"], b["v_in"], -1)
            if has_right:
                self.e.set_boundary(capi.BND_REMOTE_G_TOP, self._ptr(b["v_in"]))
        self.e.solve_phase(1)
        if self.world > 1:
            if has_right:
                self.e.get_boundary(capi.BND_X_BOT, self._ptr(b["v_out"]))
            self._sync(bvec)
            self._shift(b["v_out"], b["v_in"], +1)
            if has_left:
                self.e.set_boundary(capi.BND_REMOTE_X_BOT, self._ptr(b["v_in"]))
        self.e.solve_phase(2)
        return xvec

    def set_scaling(self, rscale, cscale):
        """Equilibrate this rank's rows (spk_set_scaling): rscale / cscale are the scales of the rank's own rows and
        columns (tensors on its device); the kp column scales of either neighbour that the halo tiles need travel
        over the process group once."""
        import torch
        if self._bufs is None:
            self._alloc(cscale)
        kp = self.kp
        from_right = torch.ones(kp, dtype=cscale.dtype, device=cscale.device)
        from_left = torch.ones(kp, dtype=cscale.dtype, device=cscale.device)
        if self.world > 1:
            self._shift(cscale[:kp].contiguous(), from_right, -1)   # my first kp go left; I get the right rank's first kp
            self._shift(cscale[-kp:].contiguous(), from_left, +1)   # my last kp go right; I get the left rank's last kp
        ext = torch.cat([from_left, cscale, from_right]).contiguous() if self.world > 1 else cscale.contiguous()
        self._scales = (rscale.contiguous(), ext)                    # keep them alive until the copy has been enqueued
        self.e.set_scaling(self._ptr(self._scales[0]), self._ptr(ext))

    def mult(self, xvec, yvec):
        """y = A x on this rank's rows with the kp-entry halos of both neighbours."""
        if self._bufs is None:
            self._alloc(xvec)
        kp = self.kp
        if self.world > 1:
            import torch
            lo = xvec[:kp].contiguous()
            hi = xvec[-kp:].contiguous()
            from_right = torch.zeros_like(lo)
            from_left = torch.zeros_like(hi)
            self._sync(xvec)
            self._shift(lo, from_right, -1)   # my first entries go left; I receive the right rank's first entries
            self._shift(hi, from_left, +1)    # my last entries go right; I receive the left rank's last entries
            if self.rank + 1 < self.world:
                self.e.set_boundary(capi.BND_HALO_RIGHT, self._ptr(from_right))
            if self.rank > 0:
                self.e.set_boundary(capi.BND_HALO_LEFT, self._ptr(from_left))
        self.e.mult(self._ptr(xvec), self._ptr(yvec))
        return yvec

    # ---- sharded Krylov: inner KSPSolve of KSPSolve_Reorder (/root/reference/src/kspreorder.c:124) on a row-block
    #      sharded band.  Same recurrences as csrc/krylov.cu (PETSc defaults restated: x0 = 0, left preconditioning,
    #      preconditioned residual norm, ||r_k|| <= rtol ||M^-1 b||, GMRES(restart) with classical Gram-Schmidt).
    #      Operator = sharded band MatMult (kp-entry halos), preconditioner = sharded SPIKE solve (two vector
    #      exchanges), every group of inner products = ONE all-reduce of a few doubles.
    def _dots(self, pairs):
        import torch
        t = torch.stack([torch.dot(a, b) for a, b in pairs])
        if self.world > 1:
            td = self.dist
            if td is None:
                import torch.distributed as td
            td.all_reduce(t)
        return t.tolist()

    def krylov(self, bvec, xvec, method=capi.GMRES, restart=30, rtol=1e-5, maxit=10000):
        """-> (iterations, residual norm, converged); xvec receives this rank's rows of the solution."""
        import math
        import torch
        n = bvec.numel()
        new = lambda: torch.empty(n, dtype=bvec.dtype, device=bvec.device)  # noqa: E731
        x = xvec
        x.zero_()
        it, conv, res = 0, False, 0.0
        if method == capi.GMRES:
            m = restart if restart > 0 else 30
            V = [new() for _ in range(m + 1)]
            t = new()
            self.solve(bvec, V[0])
            bnorm = math.sqrt(self._dots([(V[0], V[0])])[0])
            res = bnorm
            if bnorm == 0.0:
                return 0, 0.0, True
            first = True
            while it < maxit and not conv:
                if not first:
                    self.mult(x, t)
                    torch.sub(bvec, t, out=t)
                    self.solve(t, V[0])
                    res = math.sqrt(self._dots([(V[0], V[0])])[0])
                    if res <= rtol * bnorm:
                        conv = True
                        break
                first = False
                V[0].mul_(1.0 / res)
                H = [[0.0] * m for _ in range(m + 1)]
                cs, sn, g = [0.0] * m, [0.0] * m, [0.0] * (m + 1)
                g[0] = res
                j = 0
                while j < m and it < maxit:
                    w = V[j + 1]
                    self.mult(V[j], t)
                    self.solve(t, w)
                    h = self._dots([(V[i], w) for i in range(j + 1)])      # classical Gram-Schmidt: one all-reduce
                    for i in range(j + 1):
                        w.add_(V[i], alpha=-h[i])
                    hn = math.sqrt(self._dots([(w, w)])[0])
                    for i in range(j + 1):
                        H[i][j] = h[i]
                    H[j + 1][j] = hn
                    if hn != 0.0:
                        w.mul_(1.0 / hn)
                    for i in range(j):
                        a0, a1 = H[i][j], H[i + 1][j]
                        H[i][j] = cs[i] * a0 + sn[i] * a1
                        H[i + 1][j] = -sn[i] * a0 + cs[i] * a1
                    a0, a1 = H[j][j], H[j + 1][j]
                    d = math.hypot(a0, a1)
                    cs[j], sn[j] = a0 / d, a1 / d
                    H[j][j], H[j + 1][j] = d, 0.0
                    g[j + 1] = -sn[j] * g[j]
                    g[j] = cs[j] * g[j]
                    it += 1
                    res = abs(g[j + 1])
                    j += 1
                    if res <= rtol * bnorm:
                        conv = True
                        break
                y = [0.0] * j
                for i in range(j - 1, -1, -1):
                    sacc = g[i]
                    for q in range(i + 1, j):
                        sacc -= H[i][q] * y[q]
                    y[i] = sacc / H[i][i]
                for i in range(j):
                    x.add_(V[i], alpha=y[i])
            return it, res, conv
        # BiCGStab
        r, rh, p, v, s_, t, tmp = (new() for _ in range(7))
        p.zero_(); v.zero_()
        rho = alpha = omega = 1.0
        self.solve(bvec, r)
        bnorm = math.sqrt(self._dots([(r, r)])[0])
        res = bnorm
        rh.copy_(r)
        if bnorm == 0.0:
            return 0, 0.0, True
        rho_next = self._dots([(rh, r)])[0]
        while not conv and it < maxit:
            rho1 = rho_next
            if rho1 == 0.0:
                break
            beta = (rho1 / rho) * (alpha / omega)
            p.add_(v, alpha=-omega).mul_(beta).add_(r)          # p = r + beta (p - omega v)
            self.mult(p, tmp)
            self.solve(tmp, v)
            alpha = rho1 / self._dots([(rh, v)])[0]
            torch.add(r, v, alpha=-alpha, out=s_)
            self.mult(s_, tmp)
            self.solve(tmp, t)
            tt, ts = self._dots([(t, t), (t, s_)])
            omega = 0.0 if tt == 0.0 else ts / tt
            x.add_(p, alpha=alpha).add_(s_, alpha=omega)
            torch.add(s_, t, alpha=-omega, out=r)
            rho = rho1
            it += 1
            rr, rho_next = self._dots([(r, r), (rh, r)])
            res = math.sqrt(rr)
            if res <= rtol * bnorm:
                conv = True
            if omega == 0.0:
                break
        return it, res, conv

    def _sync(self, t):
        # No host synchronisation: the engine enqueues on the legacy default stream, which is also
        # torch's current stream, and torch.distributed orders its NCCL stream against the current
        # stream with events on both sides of every p2p operation.
        return
