"""Small factor + solve cases of every kernel family, to be run under compute-sanitizer (racecheck / synccheck /
memcheck): narrow LU + sweeps + tips (kt = 13, several partitions), multi-RHS sweeps, wide LU + sweeps (K = 136), and
the peer mailbox protocol with two in-process shards.  usage: sanitize_small.py [narrow|wide|peer ...]"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import spike_petsc_b200 as sp
from spike_petsc_b200 import capi

what = sys.argv[1:] or ["narrow", "wide", "peer"]


def check(S, n, nrhs=1):
    U = torch.rand(nrhs, n, dtype=torch.float64, device='cuda'); B = torch.empty_like(U); X = torch.empty_like(U)
    for r in range(nrhs):
        S.mult(U[r].data_ptr(), B[r].data_ptr())
    S.factor()
    S.solve(B.data_ptr(), X.data_ptr(), nrhs=nrhs)
    torch.cuda.synchronize()
    err = ((X - U).norm() / U.norm()).item()
    print("rel err", err, S.view()["partitions"])
    assert err < 1e-9


if "narrow" in what:
    S = sp.Spike(partitions=3, tip_tiles=0, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(6000, 100)
    check(S, 6000); check(S, 6000, nrhs=9); S.close()
    S = sp.Spike(partitions=2, tip_tiles=-1, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(1500, 20)
    check(S, 1500); S.close()
if "wide" in what:
    S = sp.Spike(partitions=2, tip_tiles=-1, mem=sp.MEM_DEVICE); S.set_band_synthetic(2048, 136)
    U = torch.rand(2, 2048, dtype=torch.float64, device='cuda'); B = torch.empty_like(U); X = torch.empty_like(U)
    for r in range(2):
        S.mult(U[r].data_ptr(), B[r].data_ptr())
    S.factor(); S.solve(B.data_ptr(), X.data_ptr(), nrhs=2); torch.cuda.synchronize()
    print("wide rel err", ((X - U).norm() / U.norm()).item()); S.close()
if "peer" in what:
    n, k, R = 4096, 20, 2
    bounds = sp.shard_rows(n, R)
    E = []
    for r in range(R):
        e = sp.Spike(partitions=2, tip_tiles=-1, mem=sp.MEM_DEVICE, rank=r, nranks=R, row_offset=bounds[r], n_global=n)
        e.set_band_synthetic(bounds[r + 1] - bounds[r], k); E.append(e)
    ptrs = [e.peer_create()[1] for e in E]
    E[0].peer_attach(1, ptr=ptrs[1]); E[1].peer_attach(0, ptr=ptrs[0])
    for e in E: e.factor_phase(10)
    E[1].peer_post(capi.BND_WT_FIRST)
    for e in E: e.factor_phase(11)
    E[0].peer_wait(capi.BND_REMOTE_WT)
    E[0].factor_phase(1); E[0].factor_phase(2); E[1].factor_phase(1)
    bs = [torch.rand(bounds[r + 1] - bounds[r], dtype=torch.float64, device='cuda') for r in range(R)]
    xs = [torch.empty_like(b) for b in bs]
    for r in range(R): E[r].solve_phase(0, bs[r].data_ptr(), xs[r].data_ptr())
    E[1].peer_post(capi.BND_G_TOP); E[0].peer_wait(capi.BND_REMOTE_G_TOP)
    for r in range(R): E[r].solve_phase(1)
    E[0].peer_post(capi.BND_X_BOT); E[1].peer_wait(capi.BND_REMOTE_X_BOT)
    for r in range(R): E[r].solve_phase(2)
    for e in E: e.peer_check()
    print("peer ok"); [e.close() for e in E]
