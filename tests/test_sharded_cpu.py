"""N>1 host logic on CPU: world_size 2 (and 3) over gloo.  The neighbour-exchange protocol of
spike_petsc_b200.sharded.ShardedSpike is driven with a dense numpy engine that implements the same
split-phase interface as the CUDA context (one SPIKE partition per rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class NumpyEngine:
    """Reference engine with the split-phase API of capi.Spike; exchange buffers are CPU tensors."""

    def __init__(self, A, lo, hi, k, rank, world):
        self.k, self.rank, self.world = k, rank, world
        n = A.shape[0]
        self.Aloc = A[lo:hi, lo:hi].copy()
        self.B = A[hi - k:hi, hi:hi + k].copy() if hi < n else None   # coupling to the right rank
        self.C = A[lo:lo + k, lo - k:lo].copy() if lo > 0 else None   # coupling to the left rank
        self.m = hi - lo
        self.remote = {}

    def buffer_address(self, t):
        return t

    def set_scaling(self, r, c_ext):
        """spk_set_scaling of a shard: c_ext = [left neighbour's last k | my columns | right neighbour's first k]."""
        k = self.k
        r, c_ext = r.numpy(), c_ext.numpy()
        c = c_ext[k:-k] if self.world > 1 else c_ext
        self.Aloc = r[:, None] * self.Aloc * c[None, :]
        if self.B is not None:
            self.B = r[-k:, None] * self.B * c_ext[-k:][None, :]
        if self.C is not None:
            self.C = r[:k, None] * self.C * c_ext[:k][None, :]
        self.rs, self.cs = r.copy(), c.copy()

    def tip_size(self):
        return self.k

    def mult(self, x, y):
        """y = A x on this rank's rows; the neighbours' k-entry halos arrive through set_boundary."""
        k = self.k
        v = self.Aloc @ x.numpy()
        if self.C is not None:
            v[:k] += self.C @ self.remote["halo_l"]
        if self.B is not None:
            v[-k:] += self.B @ self.remote["halo_r"]
        y.copy_(torch.from_numpy(v))

    def factor_phase(self, ph):
        k, m = self.k, self.m
        if ph == 1:
            if self.B is not None:
                rhs = np.zeros((m, k)); rhs[-k:] = self.B
                self.Vb = np.linalg.solve(self.Aloc, rhs)[-k:]
            if self.C is not None:
                rhs = np.zeros((m, k)); rhs[:k] = self.C
                self.Wt = np.linalg.solve(self.Aloc, rhs)[:k]
        if ph == 2:
            self.R = np.linalg.inv(np.eye(k) - self.remote["wt"] @ self.Vb)

    def get_boundary(self, which, buf):
        from spike_petsc_b200 import capi
        src = {capi.BND_WT_FIRST: lambda: self.Wt.ravel(), capi.BND_G_TOP: lambda: self.g[:self.k],
               capi.BND_X_BOT: lambda: self.xb}[which]()
        buf.copy_(torch.from_numpy(np.ascontiguousarray(src)))

    def set_boundary(self, which, buf):
        from spike_petsc_b200 import capi
        v = buf.numpy().copy()
        if which == capi.BND_REMOTE_WT:
            self.remote["wt"] = v.reshape(self.k, self.k)
        elif which == capi.BND_REMOTE_G_TOP:
            self.remote["gt"] = v
        elif which == capi.BND_REMOTE_X_BOT:
            self.remote["xb"] = v
        elif which == capi.BND_HALO_LEFT:
            self.remote["halo_l"] = v
        elif which == capi.BND_HALO_RIGHT:
            self.remote["halo_r"] = v

    def solve_phase(self, ph, b=None, x=None):
        k = self.k
        if ph == 0:
            self.x = x
            self.g = np.linalg.solve(self.Aloc, b.numpy() * getattr(self, "rs", 1.0))
            self.rbot = None
        elif ph == 1:
            if self.B is not None:
                gb = self.g[-k:]
                t = self.remote["gt"] - self.remote["wt"] @ gb
                xt = self.R @ t
                self.xb = gb - self.Vb @ xt
                self.rbot = self.B @ xt
        elif ph == 2:
            r = np.zeros(self.m)
            if self.C is not None:
                r[:k] += self.C @ self.remote["xb"]
            if self.rbot is not None:
                r[-k:] += self.rbot
            self.x.copy_(torch.from_numpy((self.g - np.linalg.solve(self.Aloc, r)) * getattr(self, "cs", 1.0)))


class MailboxEngine(NumpyEngine):
    """The same engine with the peer-mailbox interface of the CUDA context (spk_peer_*): post = non-blocking send of
    the boundary item to the neighbour, wait = receive + set_boundary.  Exercises ShardedSpike's mailbox protocol
    (handle all-gather, phase order 10 / post / 11 / wait / 1 / 2, solve posts and waits) without a GPU."""
    peer_capable = True
    overlapped_factor = True

    def __init__(self, *a):
        super().__init__(*a)
        self.peers, self.pending = {}, []

    def factor_phase(self, ph):
        if ph == 10:
            super().factor_phase(1)          # W^(t) (and V^(b)) exist before the "LU" phase, as on the GPU
        elif ph in (11, 1):
            pass
        else:
            super().factor_phase(ph)

    def peer_create(self):
        return str(self.rank).encode().ljust(64, b" "), None

    def peer_attach(self, side, handle=None, ptr=None):
        self.peers[side] = int(handle.decode())

    def _item(self, which):
        from spike_petsc_b200 import capi
        return {capi.BND_WT_FIRST: (0, self.k * self.k), capi.BND_G_TOP: (0, self.k), capi.BND_X_BOT: (1, self.k),
                capi.BND_REMOTE_WT: (1, self.k * self.k), capi.BND_REMOTE_G_TOP: (1, self.k),
                capi.BND_REMOTE_X_BOT: (0, self.k)}[which]

    def peer_post(self, which):
        side, n = self._item(which)
        buf = torch.zeros(n, dtype=torch.float64)
        self.get_boundary(which, buf)
        self.pending.append((dist.isend(buf, self.peers[side], tag=which), buf))

    def peer_wait(self, which):
        from spike_petsc_b200 import capi
        side, n = self._item(which)
        buf = torch.zeros(n, dtype=torch.float64)
        src_tag = {capi.BND_REMOTE_WT: capi.BND_WT_FIRST, capi.BND_REMOTE_G_TOP: capi.BND_G_TOP,
                   capi.BND_REMOTE_X_BOT: capi.BND_X_BOT}[which]
        dist.recv(buf, self.peers[side], tag=src_tag)
        self.set_boundary(which, buf)

    def peer_check(self):
        for req, _ in self.pending:
            req.wait()
        self.pending = []


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, k, out, mailbox=False, scaled=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import oracle as O
        from spike_petsc_b200.sharded import ShardedSpike, shard_rows
        a = O.gen_band(n, k)
        A = np.zeros((n, n))
        for i in range(n):
            for d in range(-k, k + 1):
                if 0 <= i + d < n:
                    A[i, i + d] = a[i, d + k]
        if scaled:   # A = D1 T D2; the shards equilibrate with 1/D1, 1/D2 (halo scales exchanged by ShardedSpike)
            rng = np.random.default_rng(5)
            d1, d2 = 10.0 ** rng.uniform(-6, 6, n), 10.0 ** rng.uniform(-0.5, 0.5, n)
            A = d1[:, None] * A * d2[None, :]
        u = O.gen_vec(n, 7)
        bfull = A @ u
        bounds = shard_rows(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        eng = (MailboxEngine if mailbox else NumpyEngine)(A, lo, hi, k, rank, world)
        S = ShardedSpike(eng, rank, world)
        b = torch.from_numpy(bfull[lo:hi].copy())
        x = torch.zeros_like(b)
        if scaled:
            S.set_scaling(torch.from_numpy(1.0 / d1[lo:hi]), torch.from_numpy(1.0 / d2[lo:hi]))
        S.factor(b)
        S.solve(b, x)
        if mailbox:
            assert S._peer is True
            x.zero_()
            S.solve(b, x)                    # a second solve on the same factorisation
            S.check()
        err = float(np.abs(x.numpy() - u[lo:hi]).max())
        res = torch.tensor([err], dtype=torch.float64)
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
        if rank == 0:
            out.put(res.item())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mailbox,scaled", [(2, False, False), (3, False, False), (2, True, False), (3, True, False),
                                                  (3, False, True)])
def test_sharded_exchange_protocol_gloo(world, mailbox, scaled):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 480, 5, q, mailbox, scaled)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-11


def test_shard_rows_are_tile_aligned():
    from spike_petsc_b200.sharded import shard_rows
    for n, w in [(10_000_000, 8), (1_000_003, 4), (100, 2), (64, 8)]:
        b = shard_rows(n, w)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        assert all(v % 8 == 0 for v in b[:-1])


# ---------------------------------------------------------------------------------------------- sharded Krylov
def _krylov_worker(rank, world, port, n, k, method, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import oracle as O
        from spike_petsc_b200.sharded import ShardedSpike, shard_rows
        kb = 3                                     # preconditioner: the band of half-width kb < k of the operator
        a = O.gen_band(n, k, delta=0.6)            # not diagonally dominant: the Krylov loop has work to do
        A = np.zeros((n, n))
        for i in range(n):
            for d in range(-k, k + 1):
                if 0 <= i + d < n:
                    A[i, i + d] = a[i, d + k]
        Bm = np.triu(np.tril(A, kb), -kb)
        u = O.gen_vec(n, 7)
        bfull = A @ u
        bounds = shard_rows(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]

        class Eng(NumpyEngine):                    # SPIKE on the band Bm, MatMult with the full operator A
            def __init__(self):
                super().__init__(Bm, lo, hi, k, rank, world)
                self.Aop = A[lo:hi, max(lo - k, 0):min(hi + k, n)].copy()

            def mult(self, x, y):
                xl = self.remote["halo_l"] if lo > 0 else np.zeros(0)
                xr = self.remote["halo_r"] if hi < n else np.zeros(0)
                y.copy_(torch.from_numpy(self.Aop @ np.concatenate([xl, x.numpy(), xr])))

        S = ShardedSpike(Eng(), rank, world)
        b = torch.from_numpy(bfull[lo:hi].copy())
        x = torch.zeros_like(b)
        S.factor(b)
        its, res, conv = S.krylov(b, x, method=method, restart=30, rtol=1e-9, maxit=500)
        err = torch.tensor([float(np.abs(x.numpy() - u[lo:hi]).max())], dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            # the same solve on one process: exact band LU of Bm as preconditioner (oracle Krylov)
            band = np.zeros((n, 2 * k + 1))
            for i in range(n):
                for d in range(-kb, kb + 1):
                    if 0 <= i + d < n:
                        band[i, d + k] = A[i, i + d]
            lu, _ = O.band_lu(band)
            _, its_ref, _, rc_ref = O.krylov_band(a, lu, bfull, method=O.GMRES if method == 0 else O.BICGSTAB, restart=30, rtol=1e-9, maxit=500)
            out.put((its, conv, err.item(), its_ref, rc_ref == 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,method", [(2, 0), (3, 0), (2, 1), (3, 1)])
def test_sharded_krylov_gloo(world, method):
    """ShardedSpike.krylov (GMRES / BiCGStab with one all-reduce per group of dots, halo MatMult, sharded SPIKE
    apply) needs the same number of iterations (+-1) as the single-process oracle Krylov with the exact band solve."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_krylov_worker, args=(r, world, port, 480, 5, method, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    its, conv, err, its_ref, conv_ref = q.get(timeout=5)
    assert conv and conv_ref
    assert abs(its - its_ref) <= 1, (its, its_ref)
    assert err < 1e-6
