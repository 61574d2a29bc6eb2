"""Real processes, real CUDA IPC: the spike-tip exchange through NVLink-peer-memory mailboxes (csrc/peer.cu) between
RANKS THAT ARE SEPARATE PROCESSES.  The box the driver tests on has one GPU, so the ranks share device 0 (the mailboxes
are then peer mappings of the same device's memory, opened with cudaIpcOpenMemHandle exactly as between two GPUs, and
the kernels of the ranks time-slice); the process group that carries the 64-byte IPC handles is gloo.  What it proves
beyond the in-process tests of tests/test_gpu_sharded.py: handle export / import across address spaces, the
release/acquire flag protocol between kernels of different processes, acknowledgements, a second solve on the same
factorisation, multi-column boundary items, and the wide-band path on shards.  Parity bar: 1e-10 against the exact
band solve of the oracle (src/matbanded.c:178,190)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, k, parts, nrhs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("SPIKE_B200_PEER_TIMEOUT_S", "60")      # ranks time-slice one GPU: a wait can span context switches
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        import spike_petsc_b200 as sp
        torch.cuda.set_device(0)
        a = O.gen_band(n, k)                                       # bit-identical to the engine's generator
        U = np.stack([O.gen_vec(n, 40 + c) for c in range(nrhs)])
        Bm = np.stack([O.band_mult(a, u) for u in U])
        bounds = sp.shard_rows(n, world, k)
        lo, hi = bounds[rank], bounds[rank + 1]
        eng = sp.Spike(device=0, partitions=parts, tip_tiles=-1, mem=sp.MEM_DEVICE, rank=rank, nranks=world, row_offset=lo, n_global=n)
        eng.set_band_synthetic(hi - lo, k)
        S = sp.ShardedSpike(eng, rank, world, dist=dist, nrhs=nrhs)
        b = torch.from_numpy(np.ascontiguousarray(Bm[:, lo:hi])).cuda()
        x = torch.zeros_like(b)
        S.factor(b)
        if nrhs > 1:
            S.solve(b, x, nrhs=nrhs)
        else:
            S.solve(b[0], x[0])
        assert S._peer is True, "the CUDA IPC mailbox path was not taken"
        first = x.clone()
        x.zero_()
        if nrhs > 1:
            S.solve(b, x, nrhs=nrhs)                               # a second solve on the same factorisation
        else:
            S.solve(b[0], x[0])
        S.check()
        torch.cuda.synchronize()
        assert torch.equal(first, x)                               # bit-identical
        xs = [None] * world
        dist.all_gather_object(xs, x.cpu().numpy())
        if rank == 0:
            X = np.concatenate(xs, axis=1)
            lu, _ = O.band_lu(a)
            err = max(float(np.linalg.norm(X[c] - O.band_solve(lu, Bm[c])) / np.linalg.norm(U[c])) for c in range(nrhs))
            out.put(err)
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k,parts,nrhs", [(2, 24_000, 37, 3, 1), (3, 30_000, 100, 2, 1), (2, 20_000, 60, 2, 5), (2, 16_384, 256, 1, 3)])
def test_ranks_in_separate_processes_exchange_through_ipc_mailboxes(world, n, k, parts, nrhs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, parts, nrhs, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-10
