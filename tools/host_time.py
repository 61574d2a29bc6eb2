"""Host time per factor + solve step (enqueue only) against the device time of the same steps: is a small configuration
launch bound on the host?  usage: host_time.py [n,k,P,tip ...]"""
import os, sys, time; sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
cases = [(100_000, 10, 592, 12), (1_000_000, 50, 592, 48), (1_250_000, 100, 296, 78)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
K = 200
for n, k, P, tip in cases:
    S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE, timing=False); S.keep_original(True); S.set_band_synthetic(n, k)
    u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
    S.mult(u.data_ptr(), b.data_ptr())
    for _ in range(5):
        S.factor(); S.solve(b.data_ptr(), x.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        S.factor(); S.solve(b.data_ptr(), x.data_ptr())
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"n={n} k={k} P={P} tip={tip}: host enqueue {1e3 * (t1 - t0) / K:.4f} ms per step, device {e0.elapsed_time(e1) / K:.4f} ms per step", flush=True)
    S.close()
