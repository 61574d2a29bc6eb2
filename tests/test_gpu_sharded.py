"""Sharded (multi-GPU) path on ONE GPU: R row-block shards are R contexts on the same device and the
boundary buffers are moved by plain tensor copies in the order ShardedSpike uses with NCCL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(spk, oracle, n, k, R, parts, tip, delta=1.2, overlapped=False, mailbox=False):
    import torch
    from spike_petsc_b200 import capi
    bounds = spk.shard_rows(n, R)
    E = []
    for r in range(R):
        e = spk.Spike(partitions=parts, tip_tiles=tip, mem=spk.MEM_DEVICE, rank=r, nranks=R,
                      row_offset=bounds[r], n_global=n)
        e.set_band_synthetic(bounds[r + 1] - bounds[r], k, delta=delta)
        E.append(e)
    kp = E[0].tip_size()
    dev = "cuda"
    a = oracle.gen_band(n, k, delta=delta)
    u = oracle.gen_vec(n, 3)
    bfull = oracle.band_mult(a, u)
    # ---- sharded MatMult with halos reproduces b
    xs = [torch.from_numpy(u[bounds[r]:bounds[r + 1]].copy()).to(dev) for r in range(R)]
    ys = [torch.empty_like(x) for x in xs]
    for r in range(R):
        if r > 0:
            E[r].set_boundary(capi.BND_HALO_LEFT, xs[r - 1][-kp:].contiguous().data_ptr())
        if r + 1 < R:
            E[r].set_boundary(capi.BND_HALO_RIGHT, xs[r + 1][:kp].contiguous().data_ptr())
        E[r].mult(xs[r].data_ptr(), ys[r].data_ptr())
    torch.cuda.synchronize()
    y = np.concatenate([t.cpu().numpy() for t in ys])
    assert np.linalg.norm(y - bfull) / np.linalg.norm(bfull) < 1e-14
    # ---- factor with the W^(t) exchange
    wt = [torch.zeros(kp * kp, dtype=torch.float64, device=dev) for _ in range(R)]
    if mailbox:
        # NVLink peer mailboxes (csrc/peer.cu); the shards share one stream here, so every post is queued before
        # the wait that consumes it -- the order the kernels of R processes would reach by spinning
        ptrs = [e.peer_create()[1] for e in E]
        for r in range(R):
            if r > 0:
                E[r].peer_attach(0, ptr=ptrs[r - 1])
            if r + 1 < R:
                E[r].peer_attach(1, ptr=ptrs[r + 1])
        for e in E:
            e.factor_phase(10)
        for r in range(1, R):
            E[r].peer_post(capi.BND_WT_FIRST)
        for e in E:
            e.factor_phase(11)
        for r in range(R):
            if r + 1 < R:
                E[r].peer_wait(capi.BND_REMOTE_WT)
            E[r].factor_phase(1)
            if r + 1 < R:
                E[r].factor_phase(2)
    elif not overlapped:
        for e in E:
            e.factor_phase(0); e.factor_phase(1)
        for r in range(1, R):
            E[r].get_boundary(capi.BND_WT_FIRST, wt[r].data_ptr())
        for r in range(R - 1):
            E[r].set_boundary(capi.BND_REMOTE_WT, wt[r + 1].data_ptr())
            E[r].factor_phase(2)
    else:   # the order ShardedSpike uses with NCCL: first W^(t) before the LU, boundary block with the local ones
        for e in E:
            e.factor_phase(10)
        for r in range(1, R):
            E[r].get_boundary(capi.BND_WT_FIRST, wt[r].data_ptr())
        for e in E:
            e.factor_phase(11)
        for r in range(R):
            if r + 1 < R:
                E[r].set_boundary(capi.BND_REMOTE_WT, wt[r + 1].data_ptr())
            E[r].factor_phase(1)
            if r + 1 < R:
                E[r].factor_phase(2)
    # ---- solve with the two vector exchanges
    bs = [torch.from_numpy(bfull[bounds[r]:bounds[r + 1]].copy()).to(dev) for r in range(R)]
    xo = [torch.empty_like(b) for b in bs]
    v = [torch.zeros(kp, dtype=torch.float64, device=dev) for _ in range(R)]
    for rep in range(3 if mailbox else 0):   # three solves: both slots of every channel and the ack back-pressure
        for r in range(R):
            xo[r].zero_()
            E[r].solve_phase(0, bs[r].data_ptr(), xo[r].data_ptr())
        for r in range(1, R):
            E[r].peer_post(capi.BND_G_TOP)
        for r in range(R - 1):
            E[r].peer_wait(capi.BND_REMOTE_G_TOP)
        for r in range(R):
            E[r].solve_phase(1)
        for r in range(R - 1):
            E[r].peer_post(capi.BND_X_BOT)
        for r in range(1, R):
            E[r].peer_wait(capi.BND_REMOTE_X_BOT)
        for r in range(R):
            E[r].solve_phase(2)
        for e in E:
            e.peer_check()
    for r in range(R if not mailbox else 0):
        E[r].solve_phase(0, bs[r].data_ptr(), xo[r].data_ptr())
    if not mailbox:
        for r in range(1, R):
            E[r].get_boundary(capi.BND_G_TOP, v[r].data_ptr())
        for r in range(R - 1):
            E[r].set_boundary(capi.BND_REMOTE_G_TOP, v[r + 1].data_ptr())
        for r in range(R):
            E[r].solve_phase(1)
        for r in range(R - 1):
            E[r].get_boundary(capi.BND_X_BOT, v[r].data_ptr())
        for r in range(1, R):
            E[r].set_boundary(capi.BND_REMOTE_X_BOT, v[r - 1].data_ptr())
        for r in range(R):
            E[r].solve_phase(2)
    torch.cuda.synchronize()
    x = np.concatenate([t.cpu().numpy() for t in xo])
    lu, _ = oracle.band_lu(a)
    xref = oracle.band_solve(lu, bfull)
    for e in E:
        e.close()
    return np.linalg.norm(x - xref) / np.linalg.norm(xref)


@pytest.mark.parametrize("n,k,R,parts,tip", [(40_000, 20, 2, 4, -1), (64_000, 100, 4, 3, -1), (64_000, 100, 4, 3, 0),
                                              (30_008, 37, 3, 1, -1), (200_000, 50, 8, 2, 0)])
def test_sharded_matches_reference_cpu_path(spk, oracle, n, k, R, parts, tip):
    assert _run(spk, oracle, n, k, R, parts, tip) < 1e-10


@pytest.mark.parametrize("n,k,R,parts,tip", [(64_000, 100, 4, 3, 0), (30_008, 37, 3, 1, -1), (200_000, 50, 8, 2, 0)])
def test_sharded_overlapped_factor_protocol(spk, oracle, n, k, R, parts, tip):
    """factor phases 10/11: the W^(t) exchange overlaps the band LU, the boundary block rides with the local ones."""
    assert _run(spk, oracle, n, k, R, parts, tip, overlapped=True) < 1e-10


@pytest.mark.parametrize("n,k,R,parts,tip", [(64_000, 100, 4, 3, 0), (30_008, 37, 3, 1, -1), (200_000, 50, 8, 2, 0)])
def test_sharded_peer_mailbox_protocol(spk, oracle, n, k, R, parts, tip):
    """spk_peer_post / spk_peer_wait: the boundary items travel through the neighbours' mailboxes (flag + ack words)."""
    assert _run(spk, oracle, n, k, R, parts, tip, mailbox=True) < 1e-10


def test_peer_mailbox_spin_is_bounded(spk, monkeypatch):
    """A wait whose item never arrives gives up after SPIKE_B200_PEER_TIMEOUT_S seconds (no hung GPU) and the failure
    is loud: the destination is filled with NaN, nothing is acknowledged, spk_peer_check reports it, and the next
    factor / solve call on the context fails."""
    import torch
    from spike_petsc_b200 import capi
    monkeypatch.setenv("SPIKE_B200_PEER_TIMEOUT_S", "1")
    E = [spk.Spike(partitions=2, mem=spk.MEM_DEVICE, rank=r, nranks=2, row_offset=r * 8000, n_global=16000) for r in range(2)]
    for e in E:
        e.set_band_synthetic(8000, 10)
    ptrs = [e.peer_create()[1] for e in E]
    E[0].peer_attach(1, ptr=ptrs[1])
    E[1].peer_attach(0, ptr=ptrs[0])
    with pytest.raises(RuntimeError):
        E[0].peer_post(capi.BND_REMOTE_WT)          # not an "out" item
    E[0].factor_phase(10)
    E[0].factor_phase(11)
    E[0].peer_wait(capi.BND_REMOTE_WT)              # nobody posted
    E[0].factor_phase(1)                            # enqueues the mirror copy of the error word
    E[0].factor_phase(2)
    with pytest.raises(RuntimeError, match="timed out"):
        E[0].peer_check()
    kp = E[0].tip_size()
    buf = torch.zeros(kp * kp, dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="timed out"):
        E[0].factor_phase(10)                       # the context stays failed until the mailbox is recreated
    del buf
    for e in E:
        e.close()


def test_sharded_equilibration(spk, oracle):
    """spk_set_scaling on a sharded band: every shard passes its own column scales framed by the neighbours' kp halo
    scales; the solve returns the solution of the ORIGINAL system (A = D1 T D2, rows over sixteen decades)."""
    import torch
    from spike_petsc_b200 import capi
    n, k, R, parts = 24_000, 20, 3, 2
    t = oracle.gen_band(n, k)
    rng = np.random.default_rng(11)
    d1, d2 = 10.0 ** rng.uniform(-8, 8, n), 10.0 ** rng.uniform(-0.5, 0.5, n)
    a = np.zeros_like(t)
    for d in range(-k, k + 1):
        lo, hi = max(0, -d), min(n, n - d)
        a[lo:hi, d + k] = d1[lo:hi] * t[lo:hi, d + k] * d2[lo + d:hi + d]
    u = oracle.gen_vec(n, 9)
    b = oracle.band_mult(a, u)
    lu, _ = oracle.band_lu(a)
    xref = oracle.band_solve(lu, b)
    bounds = spk.shard_rows(n, R)
    E = []
    for r in range(R):
        lo, hi = bounds[r], bounds[r + 1]
        e = spk.Spike(partitions=parts, mem=spk.MEM_DEVICE, rank=r, nranks=R, row_offset=lo, n_global=n)
        # this shard's rows of the band; the columns reaching into the neighbours stay where they are in ROWS layout
        rows = torch.from_numpy(np.ascontiguousarray(a[lo:hi])).cuda()
        e.set_band_dense_device(rows.data_ptr(), hi - lo, k)
        E.append(e)
    kp = E[0].tip_size()
    rs, cs = 1.0 / d1, 1.0 / d2
    keep = []
    for r in range(R):
        lo, hi = bounds[r], bounds[r + 1]
        left = cs[lo - kp:lo] if r > 0 else np.ones(kp)
        right = cs[hi:hi + kp] if r + 1 < R else np.ones(kp)
        rt = torch.from_numpy(rs[lo:hi].copy()).cuda()
        ct = torch.from_numpy(np.concatenate([left, cs[lo:hi], right])).cuda()
        keep += [rt, ct]
        E[r].set_scaling(rt.data_ptr(), ct.data_ptr())
    wt = [torch.zeros(kp * kp, dtype=torch.float64, device="cuda") for _ in range(R)]
    for e in E:
        e.factor_phase(0); e.factor_phase(1)
    for r in range(1, R):
        E[r].get_boundary(capi.BND_WT_FIRST, wt[r].data_ptr())
    for r in range(R - 1):
        E[r].set_boundary(capi.BND_REMOTE_WT, wt[r + 1].data_ptr())
        E[r].factor_phase(2)
    assert all(e.view()["boosted_pivots"] == 0 for e in E)
    bs = [torch.from_numpy(b[bounds[r]:bounds[r + 1]].copy()).cuda() for r in range(R)]
    xo = [torch.empty_like(v) for v in bs]
    v = [torch.zeros(kp, dtype=torch.float64, device="cuda") for _ in range(R)]
    for r in range(R):
        E[r].solve_phase(0, bs[r].data_ptr(), xo[r].data_ptr())
    for r in range(1, R):
        E[r].get_boundary(capi.BND_G_TOP, v[r].data_ptr())
    for r in range(R - 1):
        E[r].set_boundary(capi.BND_REMOTE_G_TOP, v[r + 1].data_ptr())
    for r in range(R):
        E[r].solve_phase(1)
    for r in range(R - 1):
        E[r].get_boundary(capi.BND_X_BOT, v[r].data_ptr())
    for r in range(1, R):
        E[r].set_boundary(capi.BND_REMOTE_X_BOT, v[r - 1].data_ptr())
    for r in range(R):
        E[r].solve_phase(2)
    torch.cuda.synchronize()
    x = np.concatenate([t_.cpu().numpy() for t_ in xo])
    for e in E:
        e.close()
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) < 1e-10
